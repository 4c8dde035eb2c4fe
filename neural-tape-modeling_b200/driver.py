"""Batched evaluation driver: many (file, segment) examples per launch instead of `DataLoader(batch_size=1)`.

Replaces the per-file loop of the reference's evaluation flow (code/test-model.py:332-353 prediction, :367-388 cut +
losses, :882-893 .wav output) and the segment reads of code/dataset.py:348-379 (`torchaudio.load(file, num_frames=length,
frame_offset=offset, normalize=False)`): examples are packed into one zero-padded (B, 1, T_max) batch -- the model is
causal, so trailing padding never changes a real sample --, predicted with predict() semantics per stream (zero state,
1024-sample warm start), cut by INIT_LEN, scored on the device (ESR / DCPreESR) and optionally written back as float32
.wav files.  The on-disk format is the reference's: float32 (or PCM) RIFF/WAVE, channel 0 = audio, channel 1 (if any) =
the 100 Hz pulse track.  No soundfile / torchaudio needed.
"""
import os
import struct

import numpy as np
import torch

from .losses import DCPreESR, ESRLoss

_PCM, _FLOAT, _EXTENSIBLE = 1, 3, 0xFFFE


def wav_info(path):
    """Parse the RIFF header: {'fs', 'channels', 'frames', 'format' (1 PCM / 3 float), 'bits', 'data_offset'}."""
    with open(path, "rb") as f:
        riff, _, wave = struct.unpack("<4sI4s", f.read(12))
        if riff != b"RIFF" or wave != b"WAVE":
            raise ValueError(f"{path}: not a RIFF/WAVE file")
        fmt = None
        while True:
            hdr = f.read(8)
            if len(hdr) < 8:
                raise ValueError(f"{path}: no data chunk")
            cid, size = struct.unpack("<4sI", hdr)
            if cid == b"fmt ":
                raw = f.read(size)
                tag, ch, fs, _, align, bits = struct.unpack("<HHIIHH", raw[:16])
                if tag == _EXTENSIBLE and size >= 26:
                    tag = struct.unpack("<H", raw[24:26])[0]
                fmt = (tag, ch, fs, align, bits)
                if size & 1:
                    f.read(1)
            elif cid == b"data":
                if fmt is None:
                    raise ValueError(f"{path}: data chunk before fmt chunk")
                tag, ch, fs, align, bits = fmt
                if tag not in (_PCM, _FLOAT) or (tag == _FLOAT and bits not in (32, 64)) or \
                        (tag == _PCM and bits not in (8, 16, 24, 32)):
                    raise ValueError(f"{path}: unsupported sample format (tag {tag}, {bits} bits)")
                remaining = os.path.getsize(path) - f.tell()
                size = min(size, remaining) if size not in (0, 0xFFFFFFFF) else remaining
                return {"fs": fs, "channels": ch, "frames": size // align, "format": tag, "bits": bits,
                        "data_offset": f.tell()}
            else:
                f.seek(size + (size & 1), os.SEEK_CUR)


def read_wav(path, frame_offset=0, num_frames=-1):
    """-> (float32 array (channels, frames), fs).  Float files are returned as stored (the reference reads with
    normalize=False, code/dataset.py:362-365); PCM files are scaled to [-1, 1)."""
    info = wav_info(path)
    ch, bits, tag = info["channels"], info["bits"], info["format"]
    first = max(0, min(int(frame_offset), info["frames"]))
    n = info["frames"] - first if num_frames is None or num_frames < 0 else max(0, min(int(num_frames), info["frames"] - first))
    bps = bits // 8
    with open(path, "rb") as f:
        f.seek(info["data_offset"] + first * ch * bps)
        raw = f.read(n * ch * bps)
    if tag == _FLOAT:
        a = np.frombuffer(raw, dtype="<f4" if bits == 32 else "<f8").astype(np.float32, copy=False)
    elif bits == 8:
        a = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    elif bits == 16:
        a = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif bits == 32:
        a = (np.frombuffer(raw, dtype="<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
    else:                                                   # 24-bit PCM: sign-extend three little-endian bytes
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        a = ((v ^ 0x800000) - 0x800000).astype(np.float32) / 8388608.0
    return np.ascontiguousarray(a.reshape(-1, ch).T), info["fs"]


def write_wav(path, data, fs):
    """float32 IEEE .wav; data: (frames,) or (channels, frames)."""
    a = np.asarray(data, dtype=np.float32)
    a = a.reshape(1, -1) if a.ndim == 1 else a
    ch, frames = a.shape
    payload = np.ascontiguousarray(a.T).astype("<f4").tobytes()
    with open(path, "wb") as f:
        f.write(struct.pack("<4sI4s", b"RIFF", 4 + 8 + 16 + 8 + 4 + 8 + len(payload), b"WAVE"))
        f.write(struct.pack("<4sIHHIIHH", b"fmt ", 16, _FLOAT, ch, int(fs), int(fs) * ch * 4, ch * 4, 32))
        f.write(struct.pack("<4sII", b"fact", 4, frames))
        f.write(struct.pack("<4sI", b"data", len(payload)))
        f.write(payload)


class BatchedEvaluator:
    """Predict (and score) many examples per launch.

    examples: dicts with 'input_file' and optionally 'target_file', 'offset' (frames, default 0), 'length' (frames,
    default: to the end of the file) -- the fields of `AudioDataset.examples` (code/dataset.py:348-379).  `model` is an
    ntm_b200.RNN on a CUDA device; `max_streams` examples are packed per launch."""

    def __init__(self, model, max_streams=1024):
        self.model, self.max_streams = model, int(max_streams)
        self.losses = {"ESR": ESRLoss(), "DCPreESR": DCPreESR(dc_pre=True)}

    @staticmethod
    def _load(ex, key):
        a, fs = read_wav(ex[key], ex.get("offset", 0), ex.get("length", -1))
        return a[0], fs                                       # channel 0 = audio (channel 1 = pulse track)

    def run(self, examples, out_dir=None, init_len=0):
        """-> one dict per example: {'input_name', 'frames', 'fs', ['ESR', 'DCPreESR',] ['output_file']}; losses are
        computed after cutting the first `init_len` samples (INIT_LEN, code/test-model.py:323-325,367-370)."""
        dev = self.model._device()
        results = []
        with torch.inference_mode():
            for b0 in range(0, len(examples), self.max_streams):
                group = examples[b0:b0 + self.max_streams]
                loaded = [self._load(ex, "input_file") for ex in group]
                lens = [len(a) for a, _ in loaded]
                B, Tmax = len(group), max(max(lens), 1)
                xh = torch.zeros((B, 1, Tmax), dtype=torch.float32, pin_memory=True)
                for i, (a, _) in enumerate(loaded):
                    xh[i, 0, :lens[i]] = torch.from_numpy(a)
                y = self.model.predict(xh.to(dev, non_blocking=True))
                yh = y.cpu() if out_dir is not None else None
                for i, ex in enumerate(group):
                    name = os.path.basename(ex["input_file"])
                    off = ex.get("offset", 0)
                    stem, ext = os.path.splitext(name)
                    res = {"input_name": f"{stem}_[{off}:{off + lens[i]}]{ext}", "frames": lens[i], "fs": loaded[i][1]}
                    if "target_file" in ex:
                        t, _ = self._load(ex, "target_file")
                        n = min(len(t), lens[i])
                        td = torch.from_numpy(t[:n]).to(dev).reshape(1, 1, n)
                        od = y[i:i + 1, :, :n]
                        for key, fn in self.losses.items():
                            res[key] = float(fn(od[:, :, init_len:], td[:, :, init_len:]))
                    if out_dir is not None:
                        os.makedirs(out_dir, exist_ok=True)
                        res["output_file"] = os.path.join(out_dir, f"{stem}_[{off}:{off + lens[i]}]_pred.wav")
                        write_wav(res["output_file"], yh[i, 0, :lens[i]].numpy(), loaded[i][1])
                    results.append(res)
        return results
