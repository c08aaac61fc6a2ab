import json, sys
d = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/check_attn_sweep.json"))
for k, v in d.items():
    if "burst" not in v:
        print(k, v); continue
    print(k, "burst %.2f ms %.0f TF | sustained %.2f ms %.0f TF | cross %.3f ms" % (
        v["burst"]["ms"], v["burst"]["tflops"], v["sustained"]["ms"], v["sustained"]["tflops"], v["cross"]["ms"]),
        "| sp8 %.3f ms %.0f TF" % (v["sp8_shape"]["ms"], v["sp8_shape"]["tflops"]) if "sp8_shape" in v else "",
        "| acc", " ".join("%s=%.2e" % (a[4:], v[a]["rel_l2"]) for a in v if a.startswith("acc_")),
        "nan", sum(v[a]["nan"] for a in v if a.startswith("acc_")))
