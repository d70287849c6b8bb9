"""Drop-in for ``samplernn_pase.utils.SampleRNNQuantizer`` (utils.py:25-73) on CUDA kernels.

Same constructor, constants and methods; ``quantize`` returns int64 indices like the reference,
``dequantize`` returns float32.  The mu-law chain is evaluated in one pass, bit-exact with the
reference op chain as torch executes it on CUDA (see csrc/elementwise.cu).  Dequantisation is a
256(+1)-entry table lookup: the table is the reference formula (utils.py:56-57,67-73) evaluated
once at construction, so the lookup is exact by construction (including the lost sign of
``dequantize_ulaw``, SURVEY trap 2).
"""
import torch

from . import ops


class SampleRNNQuantizer:
    LINEAR_QUANT = 0
    ULAW_QUANT = 1
    _EPSILON = 1e-2
    _EPSILONs = 1e-6
    _MU = 255.
    _LOG_MU1 = 5.5451774444795623
    q_type = None
    q_levels = None

    def __init__(self, q_type_ulaw, q_levels):
        self.q_type = self.ULAW_QUANT if q_type_ulaw else self.LINEAR_QUANT
        self.q_levels = q_levels
        self._lut_cpu = self._build_table()
        self._luts = {}
        #: device counter of samples that mapped outside [0, q_levels) (reference: index error)
        self._overflow = {}

    def _build_table(self):
        if self.q_levels > 256:
            raise RuntimeError('the dequantisation table kernel holds 257 entries: q_levels must be <= 256')
        idx = torch.arange(257)                                                # the kernel always stages 257 entries
        if self.q_type == self.LINEAR_QUANT:
            return idx.float() / (self.q_levels / 2) - 1                       # utils.py:56-57
        y = idx.float() * 2.0 / self.q_levels - 1.0                            # utils.py:69
        x = (y.abs() * self._LOG_MU1).exp() - 1                                # utils.py:71
        return x.sign() * x / self._MU                                         # utils.py:72

    def lut(self, device):
        key = str(device)
        if key not in self._luts:
            self._luts[key] = self._lut_cpu.to(device)
        return self._luts[key]

    def overflow_counter(self, device):
        key = str(device)
        if key not in self._overflow:
            self._overflow[key] = torch.zeros(1, dtype=torch.int32, device=device)
        return self._overflow[key]

    def quantize_zero(self):
        return self.q_levels // 2

    def quantize(self, samples):
        return self.quantize_both(samples)[0]

    def quantize_both(self, samples, want_i64=True):
        """(int64 indices or None, uint8 indices) - the uint8 copy feeds the fused kernels."""
        if self.q_type == self.LINEAR_QUANT:
            return ops.quantize_linear(samples, want_i64=want_i64, want_u8=self.q_levels <= 256, q_levels=self.q_levels)
        if self.q_levels != 256:
            raise RuntimeError('the mu-law kernel is built for q_levels == 256 (config.default.json:40)')
        return ops.quantize_ulaw(samples, want_i64=want_i64, want_u8=True,
                                 overflow=self.overflow_counter(samples.device))

    def dequantize(self, samples):
        return ops.dequantize_lut(samples, self.lut(samples.device))

    # explicit names kept for API parity with the reference class
    def quantize_linear(self, samples):
        return ops.quantize_linear(samples, q_levels=self.q_levels)[0]

    def dequantize_linear(self, samples):
        assert self.q_type == self.LINEAR_QUANT
        return self.dequantize(samples)

    def quantize_ulaw(self, x, max_value=1.0):
        if max_value != 1.0:
            x = x / max_value
        return ops.quantize_ulaw(x)[0]

    def dequantize_ulaw(self, y):
        assert self.q_type == self.ULAW_QUANT
        return self.dequantize(y)
