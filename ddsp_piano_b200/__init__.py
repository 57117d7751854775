"""b200-ddsp-piano: the DDSP-Piano per-sample synthesis hot path (inharmonic additive
oscillator bank, filtered noise, convolution reverb) as hand-written sm_100a CUDA behind
the reference's Processor / ProcessorGroup operator API.  See DESIGN.md."""
from . import _lib
from .engine import Engine, get_engine, total_launches
from .model import PianoModel, dafx22_model, maestro_v2_model
from .processors import (DynamicSizeFilteredNoise, FeedbackDelayNetwork, InHarmonic, MultiAdd, MultiInharmonic, MultiInstrumentReverb,
                         Processor, SurrogateAdditive, ProcessorGroup, Reverb, exp_sigmoid, exp_tanh,
                         nested_lookup, polyphonic_dag)

__all__ = ['PianoModel', 'dafx22_model', 'maestro_v2_model', 'Engine', 'get_engine', 'total_launches', 'DynamicSizeFilteredNoise', 'FeedbackDelayNetwork', 'InHarmonic',
           'MultiAdd', 'MultiInharmonic', 'MultiInstrumentReverb', 'Processor', 'SurrogateAdditive', 'ProcessorGroup', 'Reverb',
           'exp_sigmoid', 'exp_tanh', 'nested_lookup', 'polyphonic_dag', '_lib']
