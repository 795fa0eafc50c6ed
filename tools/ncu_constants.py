#!/usr/bin/env python3
"""Write profiles/r02_ncu_constants.json: the per-launch figures bench.py cannot measure itself (DRAM bytes and executed
warp instructions of the dominant kernels), read from the committed-by-summary `ncu --set full` captures in gpurun_out/.
usage: python tools/ncu_constants.py"""
import csv, io, json, os, subprocess, sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    return hdr, units, data


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def kernel_entry(rep, name_part, source, **extra):
    hdr, units, data = raw(os.path.join(REPO, rep))
    for d in data:
        if name_part in d[hdr.index("Kernel Name")]:
            ir, iw, ii, it = (hdr.index(k) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "gpu__time_duration.sum"))
            e = {"kernel": d[hdr.index("Kernel Name")][:80], "dram_bytes": to_bytes(d[ir], units[ir]) + to_bytes(d[iw], units[iw]),
                 "dram_read_bytes": to_bytes(d[ir], units[ir]), "dram_write_bytes": to_bytes(d[iw], units[iw]),
                 "warp_instructions": float(d[ii].replace(",", "")), "ncu_duration": d[it] + " " + units[it], "capture": rep, "source": source}
            e.update(extra)
            return e
    raise SystemExit("no kernel matching %r in %s" % (name_part, rep))


iv = kernel_entry("gpurun_out/r2e_prof_10k_250k.ncu-rep", "scs_place_iv4_kernel", "profiles/r02_ncu_10k_250k_summary.txt", snvs=250000)
try:        # the leaf accumulators run in their own kernel from 2,048 cells on; bench.py times the pair, so the pair is counted
    leaf = kernel_entry("gpurun_out/r2e_prof_10k_250k.ncu-rep", "scs_iv4_leaf_kernel", "profiles/r02_ncu_10k_250k_summary.txt", snvs=250000)
    for k in ("dram_bytes", "dram_read_bytes", "dram_write_bytes", "warp_instructions"):
        iv[k + "_placement_kernel_alone"] = iv[k]
        iv[k] += leaf[k]
    iv["kernel"] += " + scs_iv4_leaf_kernel"
    iv["ncu_duration"] += " + " + leaf["ncu_duration"]
except SystemExit:
    pass
out = {
    "interval_kernel": iv,
    "prepass_kernel": kernel_entry("gpurun_out/r2e_prof_10k_250k.ncu-rep", "scs_permute_reduce4_kernel", "profiles/r02_ncu_10k_250k_summary.txt", snvs=250000),
}
for key, rep, part, src, scale in (("gemm_pass", "gpurun_out/prof_place_radix.ncu-rep", "scs_place_kernel", "profiles/r01_ncu_place_radix_summary.txt", 1.0),
                                   ("gemm_cache_pass1", "gpurun_out/prof_gemm_cache250.ncu-rep", "scs_place_kernel", "profiles/r01_ncu_place_cache_summary.txt", 4.0)):
    try:
        e = kernel_entry(rep, part, src)
        if scale != 1.0:            # captured at 250k SNVs (the count cache of 1M SNVs does not survive ncu's save / restore): scaled to 1M
            for k in ("dram_bytes", "dram_read_bytes", "dram_write_bytes", "warp_instructions"):
                e[k] *= scale
            e["scaled_from_snvs"] = 250000
        out[key] = e
    except (SystemExit, Exception) as exc:
        print("skipped", key, exc, file=sys.stderr)
json.dump(out, open(os.path.join(REPO, "profiles", "r02_ncu_constants.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
