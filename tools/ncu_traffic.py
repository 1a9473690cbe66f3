"""Reads the ncu CSV of one rand_svd call (tools/ncu_target_i8.py <mode> under `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum`) and writes profiles/r02_ncu_traffic.json (bench.py's roofline.traffic / roofline.step.dram_bytes) plus a
per-kernel summary.  The target generates the matrix first: kernels before the first driver kernel are skipped by name."""
import csv, json, re, sys
from collections import OrderedDict

def load(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    per = OrderedDict()
    for r in rows[1:]:
        if r[ix["ID"]] == "ID":
            continue
        k = (int(r[ix["ID"]]), r[ix["Kernel Name"]])
        d = per.setdefault(k, {})
        val = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6}.get(unit, 1)
        d[r[ix["Metric Name"]]] = val * scale
    return per

def summarise(per):
    GEN = ("lowrank", "generate", "fill_philox", "add_noise", "scale_columns", "orth_fixup_gen")
    names = [k[1] for k in per]
    # the driver call starts at the first rowmax (auto) / the first fill of Omega or GEMM after generation: take the LAST contiguous
    # block that begins with rowmax_kernel or, for fp64, with the last fill_philox
    start = 0
    for i, nm in enumerate(names):
        if "rowmax_kernel" in nm:
            start = i; break
    else:
        for i, nm in enumerate(names):
            if "fill_philox" in nm:
                start = i
    agg = OrderedDict()
    tot_b = tot_ms = 0.0
    for (idx, nm), d in list(per.items())[start:]:
        short = re.sub(r"\(.*", "", nm).replace("void ", "").replace("rnla::", "").replace("(anonymous namespace)::", "").replace("unnamed>::", "")
        b = d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
        ms = d.get("gpu__time_duration.sum", 0)
        a = agg.setdefault(short, [0, 0.0, 0.0])
        a[0] += 1; a[1] += b; a[2] += ms
        tot_b += b; tot_ms += ms
    return agg, tot_b, tot_ms, start

out = {}
text = []
for mode in ("auto", "fp64"):
    per = load(f"gpurun_out/ncu_step_{mode}.csv")
    agg, tot_b, tot_ms, start = summarise(per)
    out[f"step_dram_bytes_{mode}"] = tot_b
    out[f"step_kernel_ms_under_ncu_{mode}"] = tot_ms
    text.append(f"== one rand_svd call, 200000 x 20000, k = 100, s = 10, mode {mode}: {tot_b * 1e-9:.1f} GB of DRAM traffic (read + write), "
                f"{tot_ms:.1f} ms of kernel time under ncu (cold, serialised) = {tot_b / 128e9:.2f} x the 128 GB of algorithmic bytes")
    for k, (cnt, b, ms) in sorted(agg.items(), key=lambda kv: -kv[1][2])[:14]:
        text.append(f"   {cnt:3d} x {k[:88]:<88} {ms:8.2f} ms  {b * 1e-9:8.2f} GB  ({b / max(cnt, 1) * 1e-9:.2f} GB per launch)")
    for k, (cnt, b, ms) in agg.items():
        if mode == "auto" and "i8_mma2_kernel<0, 7" in k.replace("(bool)", "").replace("(int)", ""):
            out["dominant_kernel_dram_bytes_per_launch"] = b / cnt
    if mode == "fp64":
        big = [d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for (i, nm), d in list(per.items())[start:] if "gemm_nn_kernel" in nm]
        out["fp64_gemm_dram_bytes_per_launch"] = max(big) if big else None
json.dump(out, open("profiles/r02_ncu_traffic.json", "w"), indent=1)
open("profiles/r02_ncu_step_summary.txt", "w").write("\n".join(text) + "\n")
print("\n".join(text)); print(out)
