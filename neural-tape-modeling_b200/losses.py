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
        sums = lib.ops().esr_sums(o, t, self.dc_pre)
        n = max(B * T, 1)
        return ((sums[0] / n) / (sums[1] / n + self.epsilon)).float()

    def per_example(self, output, target, first=None, count=None):
        """One loss per row of a (ragged) batch in ONE launch: row b is scored over its own window of samples
        [first[b], first[b] + count[b]) -- what the reference computes with batch size 1, file by file, after cutting
        INIT_LEN (code/test-model.py:367-370,385-397).  first / count: int64 tensors of B values or None (0 / T).
        -> float32 tensor of B losses (no host synchronisation)."""
        if not output.is_cuda or not target.is_cuda:
            raise RuntimeError("ntm_b200: the engine has no CPU path; move output/target to a CUDA device")
        if output.shape != target.shape:
            raise RuntimeError(f"ntm_b200: output {tuple(output.shape)} and target {tuple(target.shape)} differ")
        o, t = self._rows(output, "output"), self._rows(target, "target")
        B, T = o.shape
        dev = o.device
        if first is not None:
            first = torch.as_tensor(first, dtype=torch.int64).to(dev).contiguous()
        if count is not None:
            count = torch.as_tensor(count, dtype=torch.int64).to(dev).contiguous()
        sums = lib.ops().esr_sums_rows(o, t, first, count, self.dc_pre)
        f0 = torch.zeros(B, dtype=torch.int64, device=dev) if first is None else first.clamp(0, T)
        n = (T - f0) if count is None else torch.minimum(count.clamp(min=0), T - f0)
        n = n.clamp(min=1).to(torch.float64)
        return ((sums[:, 0] / n) / (sums[:, 1] / n + self.epsilon)).float()


class DCPreESR(ESRLoss):
    """`from GreyBoxDRC.loss_funcs import ESRLoss as DCPreESR` (code/test-model.py:28): dc_pre defaults to True."""

    def __init__(self, dc_pre=True):
        super().__init__(dc_pre=dc_pre)
