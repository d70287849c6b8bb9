"""samplernn_pase_b200 - B200-native (sm_100a) teacher-forced SampleRNN training step behind the
``samplernn_pase`` module API.  See DESIGN.md / INTEGRATION.md.

Importing the package does not touch CUDA; the first kernel call loads ``libsrnn_b200.so`` and
raises if it is missing (there is no CPU or eager fallback)."""
from . import _lib, ops, synthetic                                             # noqa: F401
from .model import CondsMixer, FrameLevelLayer, SampleLevelLayer, SampleRNNModel  # noqa: F401
from .optimizer import AdamClipped                                             # noqa: F401
from .utils import SampleRNNQuantizer                                          # noqa: F401

__all__ = ['SampleRNNModel', 'FrameLevelLayer', 'SampleLevelLayer', 'CondsMixer', 'SampleRNNQuantizer', 'AdamClipped']
