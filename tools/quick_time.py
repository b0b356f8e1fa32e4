import sys, time, hashlib, json
sys.path.insert(0, "tests")
import numpy as np, dsvlibs as L
gpu = L.gpu()
gold = json.load(open("tests/golden/streams.json"))
for name in ("cif_gop0_24", "hd_gop0"):
    g = gold[name]
    yuv = L.synth_sequence(g["w"], g["h"], g["fmt"], g["frames"], g["seed"], g["cut"])
    cfg = L.make_cfg(g["w"], g["h"], g["fmt"], gop=g["gop"], qp=g["qp"])
    for rep in range(2):
        t = time.time(); s, pk, sec = gpu.encode_sequence(cfg, yuv, g["frames"]); dt = time.time() - t
        print(name, "md5 ok" if hashlib.md5(s).hexdigest() == g["dsv_md5"] else "MD5 MISMATCH", len(s), g["dsv_len"],
              "enc fps %.1f (in-call %.1f)" % (g["frames"] / dt, g["frames"] / sec), flush=True)
if L.have_ref():
    ref = L.ref()
    g = gold["hd_gop0"]
    yuv = L.synth_sequence(g["w"], g["h"], g["fmt"], 6, g["seed"], g["cut"])
    cfg = L.make_cfg(g["w"], g["h"], g["fmt"], gop=g["gop"], qp=g["qp"])
    s, pk, sec = ref.encode_sequence(cfg, yuv, 6)
    print("ref hd_gop0 enc fps %.2f" % (6 / sec))
