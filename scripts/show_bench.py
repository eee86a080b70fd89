import json, sys
d = json.load(open(sys.argv[1]))
print("videos/s %.1f  ms/step %.3f  e2e %.1f  launches %d  sum_kernel_ms %.3f  clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["sum_kernel_ms_per_step"], d["clocks"]))
for k in d["kernels"][: int(sys.argv[2]) if len(sys.argv) > 2 else 100]:
    ms = k["ms_per_launch"] * k["launches_per_step"]
    print("%-32s %.3f ms/step (x%.0f)  %s %.0f (%.3f) tflops %.1f gbs %.0f" % (k["name"], ms, k["launches_per_step"], k["bound"], k["achieved"], k["frac"], k.get("tflops", 0), k.get("gbs", 0)))
