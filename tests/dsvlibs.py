"""ctypes loaders shared by tests/, bench.py and __graft_entry__.py.

Three libraries export the same flat signatures (see oracle/ref_harness.c):
  ref   oracle/_ref/libdsv1ref.so   the unmodified reference (checker / CPU baseline only)
  port  oracle/_ref/libdsv1port.so  our plain-C restatement  (checker only)
  gpu   digital-subband-video-1_b200/libdsv1_b200.so  the product (CUDA, sm_100a)
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "digital-subband-video-1_b200")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libdsv1ref.so")
PORT_SO = os.path.join(ROOT, "oracle", "_ref", "libdsv1port.so")
GPU_SO = os.environ.get("DSV1_B200_LIB", os.path.join(PKG, "libdsv1_b200.so"))  # the override is for A/B builds of one kernel
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "dsv1")

SUBSAMP = {"444": 0x0, "422": 0x4, "420": 0x5, "411": 0x8}
SHIFTS = {0x0: (0, 0), 0x4: (1, 0), 0x5: (1, 1), 0x8: (2, 0)}
MAX_QUALITY = 2047

CFG_KEYS = ["w", "h", "subsamp", "fps_num", "fps_den", "aspect_num", "aspect_den", "gop", "quality", "rc_mode",
            "bitrate", "do_scd", "scd_delta", "intra_pct", "pyr_levels", "stable_refresh", "max_q_step",
            "min_quality", "max_quality", "min_i_quality", "hm_nudge"]

u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)


def qp_to_quality(pct):
    return MAX_QUALITY * pct // 100


def make_cfg(w, h, fmt="420", gop=12, qp=85, rc_mode=0, **kw):
    """Mirror of the CLI's option handling (dsv_main.c:463-489) with CRF (rc_mode=0 in header terms)."""
    d = dict(w=w, h=h, subsamp=SUBSAMP[fmt], fps_num=30, fps_den=1, aspect_num=1, aspect_den=1, gop=gop,
             quality=qp_to_quality(qp), rc_mode=rc_mode, bitrate=0, do_scd=1, scd_delta=4, intra_pct=50,
             pyr_levels=0, stable_refresh=0, max_q_step=MAX_QUALITY // 200, min_quality=qp_to_quality(1),
             max_quality=qp_to_quality(100), min_i_quality=qp_to_quality(5), hm_nudge=1)
    d.update(kw)   # e.g. quality=..., bitrate=... (ABR: the CLI scales quality by 3/2, dsv_main.c:476-478)
    if d["stable_refresh"] == 0:
        d["stable_refresh"] = min(max(d["gop"] - 1, 1), 14)
    return (C.c_int * len(CFG_KEYS))(*[int(d[k]) for k in CFG_KEYS])


def frame_bytes(w, h, subsamp):
    hs, vs = SHIFTS[subsamp]
    return w * h + 2 * (((w + (1 << hs) - 1) >> hs) * ((h + (1 << vs) - 1) >> vs))


def plane_dims(w, h, subsamp):
    hs, vs = SHIFTS[subsamp]
    cw, ch = (w + (1 << hs) - 1) >> hs, (h + (1 << vs) - 1) >> vs
    return [(w, h), (cw, ch), (cw, ch)]


def coef_dims(w, h, subsamp):
    """dsv_mk_coefs (frame.c:29-61): chroma dims rounded up to even."""
    (yw, yh), (cw, ch), _ = plane_dims(w, h, subsamp)
    cw2, ch2 = (cw + 1) & ~1, (ch + 1) & ~1
    return [(yw, yh), (cw2, ch2), (cw2, ch2)]


def block_dims(w, h):
    """size4dim & ~7, clamped (dsv_encoder.c:556-595)."""
    def s4(d):
        return 64 if d > 1280 else 48 if d > 1024 else 32 if d > 704 else 24 if d > 352 else 16
    bw, bh = s4(w) & ~7, s4(h) & ~7
    return bw, bh, (w + bw - 1) // bw, (h + bh - 1) // bh


def ptr(a, ty=u8p):
    return a.ctypes.data_as(ty)


class Lib:
    """Thin wrapper giving the flat API numpy-friendly signatures.  prefix is ref_/port_/dsvk_."""

    def __init__(self, path, prefix, api_prefix=None, api_path=None):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        # tools/api_harness.c (a caller of the drop-in API) lives in its own test-only library next to the product's
        self.api_lib = C.CDLL(api_path) if api_path else self.lib
        self.p = prefix
        self.ap = api_prefix or prefix
        self.path = path

    def fn(self, name, api=False):
        return getattr(self.api_lib if api else self.lib, (self.ap if api else self.p) + name)

    def has(self, name, api=False):
        return hasattr(self.api_lib if api else self.lib, (self.ap if api else self.p) + name)

    # -- SBT ---------------------------------------------------------------
    def fwd_sbt(self, pix, pw, ph, cw, ch, isP):
        """pix: (ph, stride>=cw) uint8 array."""
        pix = np.ascontiguousarray(pix, dtype=np.uint8)
        assert pix.shape[0] == ph and pix.shape[1] >= cw
        out = np.zeros((ch, cw), dtype=np.int32)
        r = self.fn("fwd_sbt")(ptr(pix), C.c_int(pix.shape[1]), pw, ph, cw, ch, isP, ptr(out, i32p))
        assert r == 0, r
        return out

    def inv_sbt(self, coef, q, isP, c, pw, ph):
        coef = np.ascontiguousarray(coef, dtype=np.int32).copy()
        ch, cw = coef.shape
        out = np.zeros((ph, pw), dtype=np.uint8)
        r = self.fn("inv_sbt")(ptr(coef, i32p), cw, ch, q, isP, c, ptr(out), pw, pw, ph)
        assert r == 0, r
        return out

    # -- HZCC ----------------------------------------------------------------
    def encode_plane(self, coef, q, isP, c, stable, nbh, nbv):
        coef = np.ascontiguousarray(coef, dtype=np.int32).copy()
        ch, cw = coef.shape
        stable = np.ascontiguousarray(stable, dtype=np.uint8)
        cap = cw * ch * 8 + 64
        out = np.zeros(cap, dtype=np.uint8)
        n = self.fn("encode_plane")(ptr(coef, i32p), cw, ch, q, isP, c, ptr(stable), nbh, nbv, ptr(out), cap)
        assert n > 0, n
        return out[:n].copy(), coef

    def decode_plane(self, data, cw, ch, q, isP, c, stable, nbh, nbv):
        """data: the plane bytes *including* the leading 32-bit plen field (as written by encode_plane)."""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        plen = int.from_bytes(bytes(data[:4]), "big")
        body = np.zeros(len(data) - 4 + 64, dtype=np.uint8)   # slack: the reference reads ahead
        body[:len(data) - 4] = data[4:]
        stable = np.ascontiguousarray(stable, dtype=np.uint8)
        out = np.zeros((ch, cw), dtype=np.int32)
        r = self.fn("decode_plane")(ptr(body), plen, cw, ch, q, isP, c, ptr(stable), nbh, nbv, ptr(out, i32p))
        assert r == 0, r
        return out

    # -- HME / BMC ---------------------------------------------------------
    def hme(self, src, ref, w, h, subsamp, levels):
        bw, bh, nbh, nbv = block_dims(w, h)
        mv = np.zeros(nbh * nbv * 12, dtype=np.uint8)
        src = np.ascontiguousarray(src, dtype=np.uint8)
        ref = np.ascontiguousarray(ref, dtype=np.uint8)
        pct = self.fn("hme")(ptr(src), ptr(ref), w, h, subsamp, bw, bh, levels, ptr(mv))
        return pct, mv.view(MV_DTYPE)

    def hme_api(self, src, ref, w, h, subsamp, levels):
        """dsv_hme through its exported DSV_HME interface (tools/api_harness.c builds the pyramids): intra percentage and
        the vector fields of every level, level 0 first."""
        bw, bh, nbh, nbv = block_dims(w, h)
        mv = np.zeros((levels + 1) * nbh * nbv * 12, dtype=np.uint8)
        pct = self.fn("hme_api", api=True)(ptr(np.ascontiguousarray(src, dtype=np.uint8)), ptr(np.ascontiguousarray(ref, dtype=np.uint8)),
                                           w, h, subsamp, bw, bh, levels, ptr(mv))
        return pct, mv.view(MV_DTYPE).reshape(levels + 1, nbh * nbv)

    def pyramid(self, yuv, w, h, subsamp, levels):
        yuv = np.ascontiguousarray(yuv, dtype=np.uint8)
        out = np.zeros(w * h, dtype=np.uint8)
        ow = (C.c_int * 8)()
        oh = (C.c_int * 8)()
        r = self.fn("pyramid")(ptr(yuv), w, h, subsamp, levels, ptr(out), ow, oh)
        assert r == 0
        res, off = [], 0
        for i in range(levels):
            n = ow[i] * oh[i]
            res.append(out[off:off + n].reshape(oh[i], ow[i]).copy())
            off += n
        return res

    def sub_pred(self, mv, w, h, subsamp, inp, ref, blk=None):
        bw, bh, nbh, nbv = block_dims(w, h) if blk is None else blk
        n = frame_bytes(w, h, subsamp)
        pred = np.zeros(n, dtype=np.uint8)
        res = np.zeros(n, dtype=np.uint8)
        mvb = np.ascontiguousarray(mv).view(np.uint8)
        r = self.fn("sub_pred")(ptr(mvb), w, h, subsamp, bw, bh, ptr(np.ascontiguousarray(inp)),
                                ptr(np.ascontiguousarray(ref)), ptr(pred), ptr(res))
        assert r == 0
        return pred, res

    def add_pred(self, mv, w, h, subsamp, resid, ref, blk=None):
        bw, bh, nbh, nbv = block_dims(w, h) if blk is None else blk
        n = frame_bytes(w, h, subsamp)
        out = np.zeros(n, dtype=np.uint8)
        mvb = np.ascontiguousarray(mv).view(np.uint8)
        r = self.fn("add_pred")(ptr(mvb), w, h, subsamp, bw, bh, ptr(np.ascontiguousarray(resid)),
                                ptr(np.ascontiguousarray(ref)), ptr(out))
        assert r == 0
        return out

    # -- whole codec through the dsv_enc / dsv_dec API (tools/api_harness.c) --
    def encode_sequence(self, cfg, yuv, nframes):
        yuv = np.ascontiguousarray(yuv, dtype=np.uint8)
        cap = len(yuv) * 2 + 65536
        out = np.zeros(cap, dtype=np.uint8)
        pk = (C.c_int * (2 * nframes + 2))()
        npk = C.c_int(0)
        sec = C.c_double(0)
        f = self.fn("encode_sequence", api=True)
        f.restype = C.c_long
        n = f(cfg, ptr(yuv), nframes, ptr(out), C.c_long(cap), pk, C.byref(npk), C.byref(sec))
        assert n > 0, n
        return out[:n].tobytes(), list(pk[:npk.value]), sec.value

    def decode_stream(self, stream, w, h, subsamp, nframes, draw_info=0, to_420p=0):
        """draw_info: DSV_DECODER.draw_info (CLI -drawinfo); to_420p: the CLI's -out420p conversion of the output."""
        s = np.frombuffer(stream, dtype=np.uint8)
        cap = frame_bytes(w, h, SUBSAMP["420"] if to_420p else subsamp) * nframes
        out = np.zeros(cap, dtype=np.uint8)
        meta = (C.c_int * 7)()
        sec = C.c_double(0)
        n = self.fn("decode_stream_ex", api=True)(ptr(s), C.c_long(len(s)), ptr(out), C.c_long(cap), meta, C.byref(sec),
                                                  C.c_int(draw_info), C.c_int(to_420p))
        return n, out, list(meta), sec.value


# DSV_MV (dsv.h:137-150): union{int16 x,y | int32 all}; u8 mode, submask, lo_var, lo_tex, high_detail; 3 pad
MV_DTYPE = np.dtype([("x", "<i2"), ("y", "<i2"), ("mode", "u1"), ("submask", "u1"), ("lo_var", "u1"),
                     ("lo_tex", "u1"), ("high_detail", "u1"), ("pad", "u1", (3,))])

_cache = {}


def ref():
    if "ref" not in _cache:
        _cache["ref"] = Lib(REF_SO, "ref_")
    return _cache["ref"]


def port():
    if "port" not in _cache:
        _cache["port"] = Lib(PORT_SO, "port_")
    return _cache["port"]


def gpu():
    """The product.  No fallback: a missing/unloadable CUDA library is an error."""
    if "gpu" not in _cache:
        harness = os.path.join(os.path.dirname(GPU_SO), "libdsv1_b200_harness.so")
        if os.environ.get("DSV1_B200_LIB"):  # an A/B build of the product: the harness library binds to the default one
            harness = None
        _cache["gpu"] = Lib(GPU_SO, "dsvk_", api_prefix="dsvh_", api_path=harness)
    return _cache["gpu"]


EMU_SO = os.path.join(ROOT, "tests", "emu", "_build", "libdsv1_emu.so")


def emu():
    """TEST-ONLY CPU emulation of the CUDA kernel sources (tests/emu/cuda_emu.h); never the product."""
    if "emu" not in _cache:
        _cache["emu"] = Lib(EMU_SO, "dsvk_", api_prefix="dsvh_")
    return _cache["emu"]


def have_ref():
    return os.path.exists(REF_SO)


def synth_sequence(w, h, fmt, nframes, seed, cut=0, start=0):
    """Appendix-C content via oracle/synth.c (bit-identical to tests/synth.py)."""
    lib = port().lib
    hs, vs = SHIFTS[SUBSAMP[fmt]]
    n = frame_bytes(w, h, SUBSAMP[fmt]) * nframes
    out = np.zeros(n, dtype=np.uint8)
    lib.synth_sequence.restype = C.c_long
    got = lib.synth_sequence(w, h, hs, vs, start, nframes, seed, cut, ptr(out))
    assert got == n
    return out


def fwd_sbt_q(lib, pix, pw, ph, cw, ch, isP, c, q, stable, nbh, nbv):
    """Forward transform with the fused quantiser (product only): returns (dequantised coefs, dv side buffer)."""
    pix = np.ascontiguousarray(pix, dtype=np.uint8)
    stable = np.ascontiguousarray(stable, dtype=np.uint8)
    out = np.zeros((ch, cw), dtype=np.int32)
    dv = np.zeros(2 * (cw + ch) + 16, dtype=np.int32)
    n = lib.fn("fwd_sbt_q")(ptr(pix), C.c_int(pix.shape[1]), pw, ph, cw, ch, isP, c, q, ptr(stable), nbh, nbv,
                            ptr(out, i32p), ptr(dv, i32p))
    assert n >= 0, n
    return out, dv[:n]


# -- additive batch API (include/dsv1_b200_batch.h) ---------------------------------------------------
NSTATS = 16
STAT_KEYS = ["sbt_fwd_ms", "sbt_fwd_launches", "sbt_fwd_bytes", "sbt_inv_ms", "sbt_inv_launches", "sbt_inv_bytes",
             "kernel_launches", "h2d_bytes", "d2h_bytes", "pictures", "device", "lanes", "host_ms", "bmc_ms", "bmc_launches", "bmc_bytes"]


def _kernel_times(lib, fn, handle, reset):
    ms = (C.c_double * 64)()
    ln = (C.c_double * 64)()
    getattr(lib, fn)(handle, ms, ln, int(reset))
    lib.dsvb_kernel_name.restype = C.c_char_p
    return {lib.dsvb_kernel_name(i).decode(): {"ms": ms[i], "launches": ln[i]} for i in range(lib.dsvb_kernel_count()) if ln[i] > 0}


class BatchEncoder:
    """dsvb_enc_*: `lanes` sequences in lock step on one GPU."""

    def __init__(self, lib, cfg, lanes, device=0):
        self.lib = lib.lib
        self.lib.dsvb_enc_create.restype = C.c_void_p
        self.h = C.c_void_p(self.lib.dsvb_enc_create(cfg, lanes, device))
        assert self.h.value, "dsvb_enc_create failed"
        self.cfg = cfg

    def encode_ptrs(self, yuv_ptrs, nframes, on_device, stream_ptrs, caps):
        n = len(yuv_ptrs)
        a_y = (C.c_void_p * n)(*yuv_ptrs)
        a_s = (C.c_void_p * n)(*stream_ptrs)
        a_c = (C.c_long * n)(*caps)
        lens = (C.c_long * n)()
        rc = self.lib.dsvb_encode(self.h, n, nframes, a_y, int(on_device), a_s, a_c, lens)
        return rc, list(lens)

    def encode(self, seqs, nframes):
        """seqs: list of uint8 numpy arrays (host). Returns list of bytes."""
        outs = [np.zeros(len(s) * 2 + 65536, dtype=np.uint8) for s in seqs]
        rc, lens = self.encode_ptrs([s.ctypes.data for s in seqs], nframes, 0, [o.ctypes.data for o in outs],
                                    [len(o) for o in outs])
        assert rc == 0, rc
        return [o[:n].tobytes() for o, n in zip(outs, lens)]

    def encode_long_ptr(self, yuv_ptr, nframes, on_device, stream_ptr, cap):
        ln = C.c_long()
        info = (C.c_int * 4)()
        rc = self.lib.dsvb_encode_long(self.h, nframes, C.c_void_p(yuv_ptr), int(on_device), C.c_void_p(stream_ptr), C.c_long(cap),
                                       C.byref(ln), info)
        return rc, ln.value, list(info)

    def encode_long(self, yuv, nframes):
        """ONE sequence sharded by I-delimited chains over the lanes. Returns (bytes, info)."""
        out = np.zeros(len(yuv) * 2 + 65536, dtype=np.uint8)
        rc, ln, info = self.encode_long_ptr(yuv.ctypes.data, nframes, 0, out.ctypes.data, len(out))
        assert rc == 0, rc
        return out[:ln].tobytes(), info

    def stats(self, reset=False):
        st = (C.c_double * NSTATS)()
        self.lib.dsvb_enc_stats(self.h, st, int(reset))
        return dict(zip(STAT_KEYS, list(st)))

    def set_kernel_timing(self, on):
        self.lib.dsvb_enc_set_kernel_timing(self.h, int(on))

    def kernel_times(self, reset=False):
        """{kernel name: {ms, launches}} of every launch made by the engine's steps since the last reset"""
        return _kernel_times(self.lib, "dsvb_enc_kernel_times", self.h, reset)

    def close(self):
        if self.h:
            self.lib.dsvb_enc_destroy(self.h)
            self.h = None


class MultiGpu:
    """dsvb_multi_*: one engine + host thread per GPU behind one object (host buffers only)."""

    def __init__(self, lib, cfg, lanes, devices):
        self.lib = lib.lib
        self.lib.dsvb_multi_create.restype = C.c_void_p
        dv = (C.c_int * len(devices))(*devices)
        self.h = C.c_void_p(self.lib.dsvb_multi_create(cfg, lanes, len(devices), dv))
        assert self.h.value, "dsvb_multi_create failed"

    def encode(self, seqs, nframes):
        n = len(seqs)
        outs = [np.zeros(len(s) * 2 + 65536, dtype=np.uint8) for s in seqs]
        a_y = (C.c_void_p * n)(*[s.ctypes.data for s in seqs])
        a_s = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        a_c = (C.c_long * n)(*[len(o) for o in outs])
        lens = (C.c_long * n)()
        rc = self.lib.dsvb_multi_encode(self.h, n, nframes, a_y, a_s, a_c, lens)
        assert rc == 0, rc
        return [o[:k].tobytes() for o, k in zip(outs, lens)]

    def decode(self, streams, frame_bytes, nframes):
        n = len(streams)
        bufs = [np.frombuffer(s, dtype=np.uint8) for s in streams]
        outs = [np.zeros(frame_bytes * nframes, dtype=np.uint8) for _ in streams]
        a_s = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
        a_l = (C.c_long * n)(*[len(b) for b in bufs])
        a_o = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        a_c = (C.c_long * n)(*[len(o) for o in outs])
        fr = (C.c_int * n)()
        rc = self.lib.dsvb_multi_decode(self.h, n, a_s, a_l, a_o, a_c, fr)
        assert rc == 0, rc
        return outs, list(fr)

    def encode_long(self, yuv, nframes):
        out = np.zeros(len(yuv) * 2 + 65536, dtype=np.uint8)
        ln = C.c_long()
        info = (C.c_int * 4)()
        rc = self.lib.dsvb_multi_encode_long(self.h, nframes, C.c_void_p(yuv.ctypes.data), C.c_void_p(out.ctypes.data),
                                             C.c_long(len(out)), C.byref(ln), info)
        assert rc == 0, rc
        return out[:ln.value].tobytes(), list(info)

    def decode_long(self, stream, frame_bytes, nframes):
        buf = np.frombuffer(stream, dtype=np.uint8)
        out = np.zeros(frame_bytes * nframes, dtype=np.uint8)
        fr = C.c_int()
        rc = self.lib.dsvb_multi_decode_long(self.h, C.c_void_p(buf.ctypes.data), C.c_long(len(buf)), C.c_void_p(out.ctypes.data),
                                             C.c_long(len(out)), C.byref(fr))
        assert rc == 0, rc
        return out, fr.value

    def close(self):
        if self.h:
            self.lib.dsvb_multi_destroy(self.h)
            self.h = None


class BatchDecoder:
    def __init__(self, lib, lanes, device=0):
        self.lib = lib.lib
        self.lib.dsvb_dec_create.restype = C.c_void_p
        self.h = C.c_void_p(self.lib.dsvb_dec_create(lanes, device))
        assert self.h.value

    def decode_ptrs(self, stream_ptrs, stream_dev_ptrs, lens, out_ptrs, out_caps, out_on_device):
        n = len(stream_ptrs)
        a_s = (C.c_void_p * n)(*stream_ptrs)
        a_d = (C.c_void_p * n)(*stream_dev_ptrs) if stream_dev_ptrs else None
        a_l = (C.c_long * n)(*lens)
        a_o = (C.c_void_p * n)(*out_ptrs)
        a_c = (C.c_long * n)(*out_caps)
        fr = (C.c_int * n)()
        rc = self.lib.dsvb_decode(self.h, n, a_s, a_d, a_l, a_o, a_c, int(out_on_device), fr)
        return rc, list(fr)

    def decode_long(self, stream, frame_bytes, nframes):
        """one container, its chains spread over the lanes. Returns (pictures, count)."""
        buf = np.frombuffer(stream, dtype=np.uint8)
        out = np.zeros(frame_bytes * nframes, dtype=np.uint8)
        fr = C.c_int()
        rc = self.lib.dsvb_decode_long(self.h, C.c_void_p(buf.ctypes.data), None, C.c_long(len(buf)), C.c_void_p(out.ctypes.data),
                                       C.c_long(len(out)), 0, C.byref(fr))
        assert rc == 0, rc
        return out, fr.value

    def set_kernel_timing(self, on):
        self.lib.dsvb_dec_set_kernel_timing(self.h, int(on))

    def set_draw_info(self, mode):
        self.lib.dsvb_dec_set_draw_info(self.h, int(mode))

    def set_out420p(self, on):
        self.lib.dsvb_dec_set_out420p(self.h, int(on))

    def decode(self, streams, frame_bytes, nframes):
        bufs = [np.frombuffer(s, dtype=np.uint8) for s in streams]
        outs = [np.zeros(frame_bytes * nframes, dtype=np.uint8) for _ in streams]
        rc, fr = self.decode_ptrs([b.ctypes.data for b in bufs], None, [len(b) for b in bufs],
                                  [o.ctypes.data for o in outs], [len(o) for o in outs], 0)
        assert rc == 0, rc
        return outs, fr

    def stats(self, reset=False):
        st = (C.c_double * NSTATS)()
        self.lib.dsvb_dec_stats(self.h, st, int(reset))
        return dict(zip(STAT_KEYS, list(st)))

    def kernel_times(self, reset=False):
        return _kernel_times(self.lib, "dsvb_dec_kernel_times", self.h, reset)

    def close(self):
        if self.h:
            self.lib.dsvb_dec_destroy(self.h)
            self.h = None
