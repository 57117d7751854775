"""ddsp.effects.Reverb stand-in (ddsp v3.7.0 ddsp/effects.py, trainable=False), restated."""
import numpy as np

from . import core, processors


class Reverb(processors.Processor):
    def __init__(self, trainable=False, reverb_length=48000, add_dry=True, name='reverb'):
        super().__init__(name=name, trainable=trainable)
        if trainable:
            raise NotImplementedError('stand-in covers trainable=False only')
        self._reverb_length = reverb_length
        self._add_dry = add_dry

    def _mask_dry_ir(self, ir):
        if ir.ndim == 1:
            ir = ir[np.newaxis, :]
        if ir.ndim == 3:
            ir = ir[:, :, 0]
        return np.concatenate([np.zeros([ir.shape[0], 1], ir.dtype), ir[:, 1:]], axis=1)

    def get_controls(self, audio, ir=None):
        if ir is None:
            raise ValueError('Must provide "ir" tensor if Reverb trainable=False.')
        return {'audio': audio, 'ir': ir}

    def get_signal(self, audio, ir):
        audio, ir = core.tf_float32(audio), core.tf_float32(ir)
        ir = self._mask_dry_ir(ir)
        wet = core.fft_convolve(audio, ir, padding='same', delay_compensation=0)
        return (wet + audio) if self._add_dry else wet
