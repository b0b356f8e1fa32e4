"""Steady-state throughput of the per-picture API (dsv_enc / dsv_dec) on long clips (setup amortised):
python tools/long_time.py"""
import ctypes as C, sys, time
sys.path.insert(0, "tests")
import numpy as np, torch, dsvlibs as L
gpu = L.gpu()
for (w, h, fmt, n, gop) in [(1920, 1080, "420", 120, 12), (1920, 1080, "420", 60, 0), (3840, 2160, "444", 36, 12)]:
    sub = L.SUBSAMP[fmt]; fb = L.frame_bytes(w, h, sub)
    d = torch.empty(fb * n, dtype=torch.uint8, device="cuda")
    gpu.lib.dsvb_synth_device(w, h, sub, 0, n, 77, 0, C.c_void_p(d.data_ptr()), 0)
    yuv = d.cpu().numpy(); del d
    cfg = L.make_cfg(w, h, fmt, gop=gop)
    s, pk, sec = gpu.encode_sequence(cfg, yuv, n)
    nf, dec, meta, dsec = gpu.decode_stream(s, w, h, sub, n)
    print("%dx%d %s gop%d %d pictures: enc %.1f fps, dec %.1f fps (per-picture API, host buffers, incl. setup)" % (w, h, fmt, gop, n, n / sec, nf / dsec), flush=True)
    if L.have_ref():
        m = min(n, 13)
        rs, _, rsec = L.ref().encode_sequence(cfg, yuv[:m * fb], m)
        _, rdec, _, rdsec = L.ref().decode_stream(rs, w, h, sub, m)
        print("   reference (1 thread, %d pictures): enc %.2f fps, dec %.2f fps; streams equal: %s" % (m, m / rsec, m / rdsec, s[:len(rs) - 14] == rs[:len(rs) - 14] if gop else "n/a"))
