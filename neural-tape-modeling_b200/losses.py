"""Evaluation losses of the reference's test flow, computed on the device by the engine (csrc/esr.cu).

Drop-in for the two loss objects `code/test-model.py:250-253` builds:
    'ESR':      ESRLoss()                 code/Automated_GuitarAmpModelling/CoreAudioML/training.py:5-16
    'DCPreESR': DCPreESR(dc_pre=True)     code/GreyBoxDRC/loss_funcs.py:32-52 (DC_PreEmph :6-30 + ESR)
Same call: `loss_fcn(output, target)` with (B, 1, T) tensors -> 0-dim float32 tensor (no host synchronisation).
No CPU path: non-CUDA tensors raise.
"""
import torch

from . import lib


class ESRLoss(torch.nn.Module):
    """Error-to-signal ratio; `dc_pre=True` applies the reference's 2000-tap DC pre-emphasis filter to both signals."""

    def __init__(self, dc_pre=False):
        super().__init__()
        self.epsilon = 0.00001
        self.dc_pre = bool(dc_pre)

    @staticmethod
    def _rows(t, name):
        if t.dim() == 3:
            if t.shape[1] != 1:
                raise RuntimeError(f"ntm_b200: {name} must be (B, 1, T), got {tuple(t.shape)}")
            t = t[:, 0, :]
        elif t.dim() != 2:
            raise RuntimeError(f"ntm_b200: {name} must be (B, 1, T) or (B, T), got {tuple(t.shape)}")
        if t.dtype != torch.float32:
            t = t.float()
        if t.shape[1] > 0 and t.stride(1) != 1:
            t = t.contiguous()
        return t

    def forward(self, output, target):
        if not output.is_cuda or not target.is_cuda:
            raise RuntimeError("ntm_b200: the engine has no CPU path; move output/target to a CUDA device")
        if output.shape != target.shape:
            raise RuntimeError(f"ntm_b200: output {tuple(output.shape)} and target {tuple(target.shape)} differ")
        o, t = self._rows(output, "output"), self._rows(target, "target")
        if t.device != o.device:
            raise RuntimeError("ntm_b200: output and target live on different devices")
        B, T = o.shape
        sums = torch.empty(2, dtype=torch.float64, device=o.device)
        ldo = o.stride(0) if B > 1 else max(T, 1)
        ldt = t.stride(0) if B > 1 else max(T, 1)
        with torch.cuda.device(o.device):
            stream = torch.cuda.current_stream(o.device).cuda_stream
            lib.check(lib.load().ntm_esr_sums(o.data_ptr(), ldo, t.data_ptr(), ldt, B, T, int(self.dc_pre),
                                              sums.data_ptr(), o.device.index, stream))
        n = max(B * T, 1)
        return ((sums[0] / n) / (sums[1] / n + self.epsilon)).float()


class DCPreESR(ESRLoss):
    """`from GreyBoxDRC.loss_funcs import ESRLoss as DCPreESR` (code/test-model.py:28): dc_pre defaults to True."""

    def __init__(self, dc_pre=True):
        super().__init__(dc_pre=dc_pre)
