#!/usr/bin/env python3
"""Cut two small excerpts of reference-held audio into tests/golden/ (build container only): the first 2048 frames of one
int16 and one float32 file of results/**, with the original RIFF header and non-audio chunks kept byte for byte and only the
RIFF / data chunk sizes patched -- fixtures for tests/test_driver.py (the native reader vs scipy.io.wavfile)."""
import glob
import os
import struct

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/results"
N = 2048


def cut(src, dst):
    raw = open(src, "rb").read()
    pos, out, align = 12, bytearray(raw[:12]), None
    while pos + 8 <= len(raw):
        cid, size = struct.unpack("<4sI", raw[pos:pos + 8])
        body = raw[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            align = struct.unpack("<H", body[12:14])[0]
        if cid == b"data":
            body = body[:N * align]
            out += struct.pack("<4sI", cid, len(body)) + body
            break
        out += raw[pos:pos + 8] + body + (b"\0" if size & 1 else b"")
        pos += 8 + size + (size & 1)
    out[4:8] = struct.pack("<I", len(out) - 8)
    open(dst, "wb").write(bytes(out))


def main():
    import scipy.io.wavfile as wavfile
    files = sorted(glob.glob(os.path.join(REF, "**", "*.wav"), recursive=True))
    done = set()
    for f in files:
        _, a = wavfile.read(f)
        kind = a.dtype.name
        if kind in ("int16", "float32") and kind not in done:
            cut(f, os.path.join(ROOT, "tests", "golden", f"ref_{kind}_head.wav"))
            done.add(kind)
            print(kind, "<-", f)


if __name__ == "__main__":
    main()
