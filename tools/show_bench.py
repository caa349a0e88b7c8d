"""Print the key figures of a bench.py JSON line (developer aid for reading gpurun logs)."""
import json
import sys

l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.3g emb/s  ms/step %.4f  launches/step %s" % (l["value"], l["ms_per_step"], l.get("gpu_launches_per_step")))
r = l["roofline"]
print("gemm: %.1f TF/s frac %.3f kernel_ms %.4f share %.2f" % (r["achieved"], r["frac"], r["kernel_ms"], r["share_of_step"]))
e = l["e2e"]
print("e2e %.3g emb/s (%.4f ms)  python pipeline %.4f ms  autograd %.4f ms  serial %.4f ms" % (e["value"], e["ms_per_step"], e.get("python_pipeline", {}).get("ms_per_step", float("nan")), e["autograd_api"]["ms_per_step"], e["serial"]["ms_per_step"]))
for k, v in l.get("other_configs", {}).items():
    if k == "rowwise_hbm":
        for tag, rec in v.items():
            if isinstance(rec, dict):
                print(" ", tag, "  ".join("%s %.0f GB/s (%.2f)" % (n, x["achieved_gbs"], x["frac"]) for n, x in rec.items()))
    else:
        rf = v.get("roofline")
        print(" ", k, "ms %.4f" % v["ms"], ("roofline %.3f" % rf["frac"]) if rf else "")
k = l.get("knn")
if k:
    print("knn value %.1f q/s ms %.1f frac %.3f e2e %.1f parity %s" % (k["value"], k["ms_per_step"], k["roofline"]["frac"], k["e2e"]["value"], k["parity"]))
    for s in ("stream_scan", "stream_scan_q8"):
        print(" ", s, "%.1f q/s %.3f ms frac %.3f" % (k[s]["queries_per_sec"], k[s]["ms_per_call"], k[s]["roofline"]["frac"]))
    if "bank_mining" in k:
        m = k["bank_mining"]
        print("  mining hardest %.1f ms semihard %.1f ms (%d with candidate)" % (m["ms"], m["semihard"]["ms"], m["semihard"]["pairs_with_a_candidate"]))
        if "all_rows_hardest" in m:
            print("  all-rows hardest %.1f ms (%.3f of bf16x3 roofline); rich semihard %.1f ms (%d with candidate)" % (
                m["all_rows_hardest"]["ms"], m["all_rows_hardest"]["frac_of_bf16x3_roofline"],
                m["semihard_candidate_rich"]["ms"], m["semihard_candidate_rich"]["pairs_with_a_candidate"]))
print("clocks", l.get("clocks"))
