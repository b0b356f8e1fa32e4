"""Quick throughput probe through the drop-in API (not the bench): python tools/quick_time.py [names...]"""
import sys, time, hashlib, json
sys.path.insert(0, "tests")
import numpy as np, dsvlibs as L
gpu = L.gpu()
gold = json.load(open("tests/golden/streams.json"))
names = sys.argv[1:] or ["cif_gop12", "hd_gop0", "hd_gop12", "uhd444_gop12"]
for name in names:
    g = gold[name]
    sub = L.SUBSAMP[g["fmt"]]
    yuv = L.synth_sequence(g["w"], g["h"], g["fmt"], g["frames"], g["seed"], g["cut"])
    cfg = L.make_cfg(g["w"], g["h"], g["fmt"], gop=g["gop"], qp=g["qp"])
    for rep in range(2):
        s, pk, sec = gpu.encode_sequence(cfg, yuv, g["frames"])
        nf, dec, meta, dsec = gpu.decode_stream(s, g["w"], g["h"], sub, g["frames"])
        print(name, "enc", "ok" if hashlib.md5(s).hexdigest() == g["dsv_md5"] else "MISMATCH",
              "dec", "ok" if hashlib.md5(dec.tobytes()).hexdigest() == g["dec_md5"] else "MISMATCH",
              "enc fps %.1f dec fps %.1f" % (g["frames"] / sec, g["frames"] / dsec), flush=True)
    if L.have_ref() and "--ref" in sys.argv:
        pass
if L.have_ref():
    ref = L.ref()
    for name in names:
        g = gold[name]
        n = min(g["frames"], 13)
        yuv = L.synth_sequence(g["w"], g["h"], g["fmt"], n, g["seed"], g["cut"])
        cfg = L.make_cfg(g["w"], g["h"], g["fmt"], gop=g["gop"], qp=g["qp"])
        s, pk, sec = ref.encode_sequence(cfg, yuv, n)
        nf, dec, meta, dsec = ref.decode_stream(s, g["w"], g["h"], L.SUBSAMP[g["fmt"]], n)
        print("ref", name, "enc fps %.2f dec fps %.2f" % (n / sec, n / dsec), flush=True)
