"""ddsp.synths.FilteredNoise stand-in (ddsp v3.7.0 ddsp/synths.py), restated."""
import numpy as np

from . import core, processors


class FilteredNoise(processors.Processor):
    def __init__(self, n_samples=64000, window_size=257, scale_fn=core.exp_sigmoid,
                 initial_bias=-5.0, name='filtered_noise'):
        super().__init__(name=name)
        self.n_samples = n_samples
        self.window_size = window_size
        self.scale_fn = scale_fn
        self.initial_bias = initial_bias

    def get_controls(self, magnitudes):
        if self.scale_fn is not None:
            magnitudes = self.scale_fn(magnitudes + np.asarray(magnitudes).dtype.type(
                self.initial_bias))
        return {'magnitudes': magnitudes}
