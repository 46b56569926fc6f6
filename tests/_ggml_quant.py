"""Test helper: re-write a legacy-ggml model file with its 2-D weights block-quantised the way
whisper.cpp's quantize tool stores them (q4_0 / q4_1 / q5_0 / q5_1 / q8_0, blocks of 32 along the contiguous dimension), and
with the SAME values expanded back to f16 — so that loading either file must give identical results.  numpy only."""
import struct

import numpy as np

QK = 32
TYPES = {"q4_0": 2, "q4_1": 3, "q5_0": 6, "q5_1": 7, "q8_0": 8}
FTYPES = {"q4_0": 2, "q4_1": 3, "q8_0": 7, "q5_0": 8, "q5_1": 9}


def _quant(x: np.ndarray, kind: str):
    """x: [n_blocks, 32] f32 -> (raw bytes per block [n_blocks, bs] uint8, dequantised [n_blocks, 32] f32)."""
    n = x.shape[0]
    if kind == "q8_0":
        d = (np.abs(x).max(1) / 127.0).astype(np.float16)
        df = d.astype(np.float32)
        q = np.where(df[:, None] > 0, np.round(x / np.where(df[:, None] > 0, df[:, None], 1)), 0).clip(-127, 127).astype(np.int8)
        raw = np.concatenate([d.view(np.uint8).reshape(n, 2), q.view(np.uint8)], 1)
        return raw, q.astype(np.float32) * df[:, None]
    sym = kind in ("q4_0", "q5_0")
    bits = 4 if kind.startswith("q4") else 5
    levels = (1 << bits)
    if sym:
        idx = np.abs(x).argmax(1)
        mx = x[np.arange(n), idx]
        d = (mx / -(levels // 2)).astype(np.float16)
        df = d.astype(np.float32)
        inv = np.where(df != 0, 1.0 / np.where(df != 0, df, 1), 0.0)
        q = np.clip(np.floor(x * inv[:, None] + levels / 2 + 0.5), 0, levels - 1).astype(np.uint8)
        deq = (q.astype(np.float32) - levels // 2) * df[:, None]
        head = d.view(np.uint8).reshape(n, 2)
    else:
        mn, mx = x.min(1), x.max(1)
        d = ((mx - mn) / (levels - 1)).astype(np.float16)
        m = mn.astype(np.float16)
        df, mf = d.astype(np.float32), m.astype(np.float32)
        inv = np.where(df != 0, 1.0 / np.where(df != 0, df, 1), 0.0)
        q = np.clip(np.floor((x - mf[:, None]) * inv[:, None] + 0.5), 0, levels - 1).astype(np.uint8)
        deq = q.astype(np.float32) * df[:, None] + mf[:, None]
        head = np.concatenate([d.view(np.uint8).reshape(n, 2), m.view(np.uint8).reshape(n, 2)], 1)
    lo, hi = q[:, :16], q[:, 16:]
    qs = ((lo & 0x0F) | ((hi & 0x0F) << 4)).astype(np.uint8)
    if bits == 5:
        qh = np.zeros(n, dtype=np.uint32)
        for j in range(16):
            qh |= ((lo[:, j] >> 4) & 1).astype(np.uint32) << j
            qh |= ((hi[:, j] >> 4) & 1).astype(np.uint32) << (j + 16)
        raw = np.concatenate([head, qh.view(np.uint8).reshape(n, 4), qs], 1)
    else:
        raw = np.concatenate([head, qs], 1)
    return raw, deq


def requantise(src: str, dst_q: str, dst_f16: str, kind: str) -> int:
    """Copies `src` to `dst_q` (2-D f16 weights with rows that are a multiple of 32 become `kind` blocks) and to `dst_f16` (same
    tensors, the dequantised values rounded to f16).  Returns the number of quantised tensors."""
    data = open(src, "rb").read()
    pos = 0

    def take(n):
        nonlocal pos
        b = data[pos:pos + n]
        pos += n
        return b

    head = bytearray(take(4 + 11 * 4))
    hp = list(struct.unpack("<11i", head[4:]))
    n_mel, n_fft = struct.unpack("<2i", data[pos:pos + 8])
    filt = take(8 + n_mel * n_fft * 4)
    (n_vocab_file,) = struct.unpack("<i", data[pos:pos + 4])
    v0 = pos
    pos += 4
    for _ in range(n_vocab_file):
        (ln,) = struct.unpack("<I", data[pos:pos + 4])
        pos += 4 + ln
    vocab = data[v0:pos]
    hq = bytearray(head)
    hq[4 + 10 * 4:4 + 11 * 4] = struct.pack("<i", FTYPES[kind])
    outq, outf = [bytes(hq), filt, vocab], [bytes(head), filt, vocab]
    n_q = 0
    while pos < len(data):
        n_dims, name_len, ttype = struct.unpack("<3i", take(12))
        ne = struct.unpack(f"<{n_dims}i", take(4 * n_dims))
        name = take(name_len)
        nelem = int(np.prod(ne))
        raw = take(nelem * (2 if ttype == 1 else 4))
        hdr = lambda t: struct.pack("<3i", n_dims, name_len, t) + struct.pack(f"<{n_dims}i", *ne) + name  # noqa: E731
        if ttype == 1 and n_dims == 2 and ne[0] % QK == 0 and b"token_embedding" not in name and b"positional" not in name:
            x = np.frombuffer(raw, dtype=np.float16).astype(np.float32).reshape(-1, QK)
            rq, deq = _quant(x, kind)
            outq += [hdr(TYPES[kind]), rq.tobytes()]
            outf += [hdr(1), deq.astype(np.float16).tobytes()]
            n_q += 1
        else:
            outq += [hdr(ttype), raw]
            outf += [hdr(ttype), raw]
    open(dst_q, "wb").write(b"".join(outq))
    open(dst_f16, "wb").write(b"".join(outf))
    return n_q
