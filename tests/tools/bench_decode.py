"""VCF decode (scs_vcf_decode) parity at a larger size and throughput.  Run on the GPU box."""
import os, sys, time, tempfile
import numpy as np
sys.path.insert(0, ".")  # run from the repo root: python tests/tools/bench_decode.py
from oracle import pt_oracle as O
from scsommerclock_b200 import engine, run_PT_test as P, synth

def make_lines(n_cells, n_lines, seed, pos0=1):
    rng = np.random.default_rng(seed)
    tree = synth.coalescent_tree(n_cells, rng)
    codes = synth.observe(tree, synth.throw_mutations(tree, n_lines, rng), 0.1, 0.01, 0.15, rng)
    bases = "ACGT"
    out = []
    for i in range(n_lines):
        ref = bases[rng.integers(4)]
        alts = [b for b in bases if b != ref]
        a = 1 + int(rng.integers(3))
        recs = []
        for c in codes[i]:
            if c == 3: recs.append(".|.:4:4,0,0,0:8:0|0")
            elif c == 0: recs.append("0|0:12:12,0,0,0:10:0|0")
            else: recs.append(("%d|%d" % (a, a) if rng.random() < 0.1 else "0|%d" % a) + ":9:4,0,5,0:33:0|%d" % a)
        recs.append("0|0:20:20,0,0,0:10:0|0")
        flt = "PASS" if rng.random() > 0.1 else ("wildtype" if rng.random() < 0.5 else "s50")
        out.append("1\t%d\tsnv%06d\t%s\t%s\t.\t%s\tAA=%s;NS=%d;SOMATIC\tGT:DP:RC:GQ:TG\t%s" % (pos0 + 3 * i, i, ref, ",".join(alts), flt, ref, n_cells, "\t".join(recs)))
    return out

def main():
    n_cells, n_unique, reps = 200, 3000, 40
    names = ["cell%04d" % (i + 1) for i in range(n_cells)] + ["outgcell"]
    header = "##fileformat=VCFv4.3\n##source=synthetic\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(names) + "\n"
    lines = make_lines(n_cells, n_unique, 5)
    d = tempfile.mkdtemp()
    small = os.path.join(d, "small.vcf")
    open(small, "w").write(header + "\n".join(lines) + "\n")
    ctx = P.get_context()
    t0 = time.time(); want = O.get_mut_df(small); t_cpu = time.time() - t0
    muts, _, _ = P.get_mut_df(small, "", "", ctx=ctx)
    wc = np.where(np.isnan(want["muts"]), 3, np.nan_to_num(want["muts"])).astype(np.uint8)
    ok = np.array_equal(muts.codes, wc) and list(muts.columns) == list(want["samples"]) and muts.index == want["pos"]
    print("parity %d lines x %d samples: %s (kept %d, no-alt %d, filtered %d); python oracle %.1f s" % (n_unique, n_cells + 1, ok, muts.shape[0], muts._dec.n_noalt, muts._dec.n_filtered, t_cpu))
    big = (header + ("\n".join(lines) + "\n") * reps).encode()
    import torch
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        dec = engine.DecodedVcf(ctx, big, n_cells + 1, list(range(n_cells + 1)))
        torch.cuda.synchronize(); dt = time.time() - t0
        print("decode %.0f MB, %d lines x %d samples: %.1f ms  = %.2f GB/s of VCF text (H2D copy of the text included), %.2e genotype fields/s" % (len(big) / 1e6, dec.n_lines, n_cells + 1, dt * 1e3, len(big) / dt / 1e9, dec.n_lines * (n_cells + 1) / dt))
        dec.close()
    print("RESULT", "PASS" if ok else "FAIL")
main()
