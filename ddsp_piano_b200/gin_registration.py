"""gin bindings for the drop-in swap (reference: configs/dafx22.gin:2-6 imports the modules whose
``@gin.register`` decorators make ``@inharm_synth.MultiInharmonic()`` etc. resolvable, and :91-100 wires them
into ``processors.ProcessorGroup.dag = @polyphonic_dag()``).  Importing THIS module from a gin file

    import ddsp_piano_b200.gin_registration

registers the B200 processors under the module name ``ddsp_piano_b200`` so that the three bindings of
INTEGRATION.md section 1 select them.  ``gin-config`` is a dependency of the reference, not of this
package: without it the import fails with a message saying so (nothing else in the package needs gin).
"""
try:
    import gin
except ImportError as e:                                   # pragma: no cover - exercised with a stub in the tests
    raise ImportError('ddsp_piano_b200.gin_registration needs gin-config (pip install gin-config), the '
                      'configuration library the reference wires its models with') from e

import ddsp_piano_b200 as _dp

MODULE = 'ddsp_piano_b200'
CONFIGURABLES = ('MultiInharmonic', 'InHarmonic', 'SurrogateAdditive', 'DynamicSizeFilteredNoise', 'Reverb',
                 'MultiAdd', 'FeedbackDelayNetwork', 'MultiInstrumentReverb', 'ProcessorGroup', 'polyphonic_dag',
                 'exp_tanh', 'exp_sigmoid')

registered = {}
for _name in CONFIGURABLES:
    registered[_name] = gin.external_configurable(getattr(_dp, _name), name=_name, module=MODULE)
