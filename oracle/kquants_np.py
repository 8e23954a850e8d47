"""TEST INFRASTRUCTURE ONLY (never imported by llm_b200/): numpy restatement of the reference's K-quant hot path, small sizes only (pure-Python loops).

  quantize_row_q8_K      LC/k_quants.c:1133-1168 (quantize_row_q8_K_reference) as the reference's x86 build computes it: gcc contracts nearest_int's
                         `iscale*x + 12582912.f` into one fused multiply-add (vfmadd132ps in oracle/_ref) -- restated here with a float64 product-sum
  vec_dot                ggml_vec_dot_q{2,3,4,5,6}_K_q8_K, AVX2 branches: LC/k_quants.c:1330-1394, :1892-1996, :2593-2655, :3124-3200, :3737-3811
                         (8 int32 lanes per super-block = byte columns 4L..4L+3 of every 32-byte group weighted by the sub-block scales; one 8-lane
                         f32 fma chain over super-blocks; Q2_K mins into the same lanes first, Q4_K 4 f32 min lanes, Q5_K one scalar fma chain)
Pinned against oracle/_ref by tests/test_oracle_kquants.py; it is also the layout the CUDA kernel (llm_b200/csrc/kquants.cu) was written from."""
import numpy as np

f32 = np.float32
BLOCK_BYTES = {10: 84, 11: 110, 12: 144, 13: 176, 14: 210}


def _fma(a, b, c):
    """fmaf for the operand ranges here: the f32 x f32 product is exact in float64 and one more float64 add keeps > 24 significant bits of slack except in
    measure-zero ties, so a single rounding to f32 follows (the pinning test holds it to the reference bit for bit)."""
    return f32(np.float64(a) * np.float64(b) + np.float64(c))


def _h2f(b):
    return f32(np.frombuffer(bytes(b), np.float16)[0])


def quantize_row_q8_K(x):
    """-> list of (d: f32, qs: int32[256], bsums: int32[16]) per super-block"""
    x = np.ascontiguousarray(x, f32)
    out = []
    for i in range(x.size // 256):
        v = x[i * 256:(i + 1) * 256]
        ax = np.abs(v)
        idx = int(np.argmax(ax))                       # first element of largest magnitude (the scalar loop's strict `>`)
        if ax[idx] == 0:
            out.append((f32(0), np.zeros(256, np.int32), np.zeros(16, np.int32)))
            continue
        iscale = f32(-128.0) / f32(v[idx])
        val = (np.float64(iscale) * v.astype(np.float64) + 12582912.0).astype(f32)
        q = np.minimum((val.view(np.int32) & 0x7fffff) - 0x400000, 127).astype(np.int32)
        out.append((f32(1.0) / iscale, q, q.reshape(16, 16).sum(1)))
    return out


def q8_K_bytes(blocks):
    """the block_q8_K byte image {f32 d; i8 qs[256]; i16 bsums[16]} (LC/k_quants.h:112-117)"""
    return np.concatenate([np.concatenate([np.array([d], f32).view(np.uint8), q.astype(np.int8).view(np.uint8), bs.astype(np.int16).view(np.uint8)]) for d, q, bs in blocks])


def _k4(q, j):
    """get_scale_min_k4 (LC/k_quants.c:316-324) = the utmp shuffles of the AVX2 branches"""
    if j < 4:
        return int(q[j] & 63), int(q[j + 4] & 63)
    return int((q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4)), int((q[j + 4] >> 4) | ((q[j] >> 6) << 4))


def vec_dot(t, wrow, xq):
    """one weight row (uint8 image of K/256 super-blocks of ggml type t in 10..14) . one quantize_row_q8_K(x) -> f32"""
    nbytes = BLOCK_BYTES[t]
    acc = np.zeros(8, f32); accm = np.zeros(4, f32); summs = f32(0)
    for i, (yd, q8, bs) in enumerate(xq):
        raw = wrow[i * nbytes:(i + 1) * nbytes]
        wb = raw.astype(np.int64)
        sumi = np.zeros(8, np.int64)
        if t in (10, 11):
            if t == 10:
                d = f32(yd * _h2f(raw[80:82])); dmin = f32(f32(-yd) * _h2f(raw[82:84])); sc = wb[:16]; qs = wb[16:80]
                for L in range(8):
                    acc[L] = _fma(dmin, f32((sc[2 * L] >> 4) * bs[2 * L] + (sc[2 * L + 1] >> 4) * bs[2 * L + 1]), acc[L])
                scale = lambda b: int(sc[b] & 0xF)
                hm = None
            else:
                d = f32(yd * _h2f(raw[108:110])); hm = wb[:32]; qs = wb[32:96]; s = wb[96:108]

                def scale(b):
                    w, c = b >> 2, b & 3
                    return int(((s[(w & 1) * 4 + c] >> (4 * (w >> 1))) & 0xF) | (((s[8 + c] >> (2 * w)) & 3) << 4)) - 32
            for j in range(2):
                for k in range(4):
                    for L in range(8):
                        dot = 0
                        for e in range(4):
                            q = (qs[32 * j + 4 * L + e] >> (2 * k)) & 3
                            if hm is not None and not ((hm[4 * L + e] >> (4 * j + k)) & 1):
                                q -= 4
                            dot += int(q) * int(q8[128 * j + 32 * k + 4 * L + e])
                        sumi[L] += scale(8 * j + 2 * k + (L >= 4)) * dot
        elif t == 14:
            d = f32(yd * _h2f(raw[208:210])); ql = wb[:128]; qh = wb[128:192]; sc = raw[192:208].view(np.int8).astype(np.int64)
            for j in range(2):
                for k in range(4):
                    for L in range(8):
                        dot = 0
                        for e in range(4):
                            bl = ql[64 * j + 32 * (k & 1) + 4 * L + e]
                            q = ((bl >> 4) if k >= 2 else (bl & 15)) | (((qh[32 * j + 4 * L + e] >> (2 * k)) & 3) << 4)
                            dot += (int(q) - 32) * int(q8[128 * j + 32 * k + 4 * L + e])
                        sumi[L] += int(sc[2 * (4 * j + k) + (L >= 4)]) * dot
        else:
            d = f32(yd * _h2f(raw[0:2])); dmin = f32(f32(-yd) * _h2f(raw[2:4])); scq = wb[4:16]
            QS = 16 if t == 12 else 48
            for g in range(8):                           # 32-byte group g: nibble g & 1 of qs[32 (g >> 1) ..] against q8[32 g ..]
                for L in range(8):
                    dot = 0
                    for e in range(4):
                        b = wb[QS + 32 * (g >> 1) + 4 * L + e]
                        q = (b >> 4) if g & 1 else (b & 15)
                        if t == 13:
                            q |= ((wb[16 + 4 * L + e] >> g) & 1) << 4
                        dot += int(q) * int(q8[32 * g + 4 * L + e])
                    sumi[L] += _k4(scq, g)[0] * dot
            prod = [_k4(scq, 2 * tt)[1] * int(bs[4 * tt] + bs[4 * tt + 1]) + _k4(scq, 2 * tt + 1)[1] * int(bs[4 * tt + 2] + bs[4 * tt + 3]) for tt in range(4)]
            if t == 12:
                for tt in range(4):
                    accm[tt] = _fma(dmin, f32(prod[tt]), accm[tt])
            else:
                summs = _fma(dmin, f32(sum(prod)), summs)
        for L in range(8):
            acc[L] = _fma(d, f32(sumi[L]), acc[L])
    v = f32(f32(f32(acc[0] + acc[4]) + f32(acc[2] + acc[6])) + f32(f32(acc[1] + acc[5]) + f32(acc[3] + acc[7])))     # hsum_float_8, LC/k_quants.c:1193-1199
    if t == 12:
        v = f32(v + f32(f32(accm[0] + accm[2]) + f32(accm[1] + accm[3])))
    if t == 13:
        v = f32(v + summs)
    return v
