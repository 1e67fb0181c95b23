import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import test_gpu_bench_modes as T
from oracle import Oracle
from raytrace3_b200 import Context, abi
from raytrace3_b200.scene import reference_camera
ctx = Context(0); orc = Oracle()
def margins(scene_name, w, h, count, pass_id, begin=0):
    sc = T._scene(scene_name)
    ctx.set_params(T.bench_params()); ctx.upload_scene(sc)
    cam = reference_camera(w, h)
    ref_e = orc.eye_pass(sc, T.exact_params(), cam, pass_id=0)
    n = ctx.eye_pass(cam, 0); got = ctx.export_hitpoints()
    kr = (ref_e["pix"][:, 0].astype(np.int64) * w + ref_e["pix"][:, 1]) * (1 << 20) + ref_e["node"]
    kg = (got["pix"][:, 0].astype(np.int64) * w + got["pix"][:, 1]) * (1 << 20) + got["node"]
    ig, ir = T._match(kg, kr)
    rel = np.abs(got["pos"][ig] - ref_e["pos"][ir]) / np.maximum(1.0, np.abs(ref_e["pos"][ir]))
    print(scene_name, "eye: common", len(ig), "of", n, len(kr), "max rel", rel.max(), "n err", np.abs(got["n"][ig]-ref_e["n"][ir]).max())
    ctx.build_index(); hp = ctx.export_hitpoints()
    ref = orc.photon_pass(sc, T.exact_params(), begin, count, pass_id, hp=hp, keep_deposits=True)
    ctx.keep_deposits(True)
    ctx.reset_counters(); ctx.photon_pass(begin, count, pass_id); ctx.update(); dep = ctx.export_deposits(); c = ctx.counters()
    ctx.keep_deposits(False)
    rc = ref["counters"]; rd = ref["dep"]
    print("  counters: queries", c["isect_queries"], rc["queries"], "deposits", c["deposits"], rc["deposits"], "solves", c["patch_solves"], rc["solves"])
    kg = dep["photon"].astype(np.int64) * 64 + dep["bounce"]; kr = rd["photon"].astype(np.int64) * 64 + rd["bounce"]
    ig, ir = T._match(kg, kr)
    rel = (np.abs(dep["x"][ig] - rd["x"][ir]) / np.maximum(1.0, np.abs(rd["x"][ir]))).max(1)
    same = rel <= 1e-6
    fg, fr = dep["flux"][ig][same], rd["flux"][ir][same]
    print("  deposits same frac", same.sum() / len(kr), "(bar 0.995)  flux close frac", np.isclose(fg, fr, rtol=1e-6, atol=1e-12).mean(), "(bar 0.9995) max flux rel", (np.abs(fg-fr)/np.maximum(np.abs(fr),1e-12)).max(), "(bar 2e-3)")
    ctx.build_index(); ctx.photon_pass(begin, count, pass_id); ctx.update()
    g = ctx.export_stats(hp["pos"].shape[0])
    print("  newN equal frac", (g["newN"] == ref["newN"]).mean(), "(bar 0.995) sum diff rel", abs(int(g["newN"].sum()) - int(ref["newN"].sum())) / ref["newN"].sum(), "(bar 0.002) tao sum rel", np.abs(g["tao"].sum(0)/ref["tao"].sum(0)-1).max(), "(bar 2e-3)")
margins("reference", 320, 240, 100_000, 1)
margins("cornell", 256, 192, 60_000, 2, begin=1000)
margins("stress", 192, 108, 40_000, 3)
