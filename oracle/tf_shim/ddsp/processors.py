"""ddsp.processors stand-in: Processor.__call__ and ProcessorGroup/DAGLayer.run_dag
semantics of ddsp v3.7.0 (ddsp/processors.py, ddsp/dags.py), restated."""
from . import core


class Processor:
    def __init__(self, name, trainable=True):
        self.name = name
        self.trainable = trainable

    def __call__(self, *args, return_outputs_dict=False, **kwargs):
        for k in ['training', 'mask']:
            kwargs.pop(k, None)
        controls = self.get_controls(*args, **kwargs)
        signal = self.get_signal(**controls)
        if return_outputs_dict:
            return dict(signal=signal, controls=controls)
        return signal

    def get_controls(self, *args, **kwargs):
        raise NotImplementedError

    def get_signal(self, *args, **kwargs):
        raise NotImplementedError

    def build(self, input_shape=None):     # keras Layer.build: nothing to do in the stand-in
        self.built = True


class ProcessorGroup:
    """DAG of (processor, [input keys]) nodes run in order over a growing outputs dict."""

    def __init__(self, dag, name='processor_group'):
        self.name = name
        self.dag = []
        self._modules = {}
        for node in dag:
            module, keys = node[0], node[1]
            self._modules[module.name] = module
            self.dag.append((module.name, keys))

    @property
    def processors(self):
        return [self._modules[k] for k, _ in self.dag]

    def get_controls(self, inputs, **kwargs):
        outputs = inputs
        module_outputs = None
        for module_key, input_keys in self.dag:
            module = self._modules[module_key]
            args = [core.nested_lookup(key, outputs) for key in input_keys]
            module_outputs = module(*args, return_outputs_dict=True, **kwargs)
            outputs[module_key] = module_outputs
        outputs['out'] = module_outputs
        return outputs

    def get_signal(self, outputs):
        return outputs['out']['signal']

    def __call__(self, inputs, return_outputs_dict=False, **kwargs):
        controls = self.get_controls(inputs, **kwargs)
        signal = self.get_signal(controls)
        if return_outputs_dict:
            return dict(signal=signal, controls=controls)
        return signal
