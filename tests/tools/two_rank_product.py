#!/usr/bin/env python3
"""Runs the product path under an initialised process group (launched by torchrun from
tests/test_gpu_parity.py::test_two_ranks_*): first ONE dataset with its VCF record lines sharded over the
ranks (run_poisson_tree_test: kept-SNV counts, missing counts and soft weights all-reduced, rank 0 writes
the TSV), then a replicate sweep with whole (VCF, tree) pairs dealt to the ranks
(run_poisson_tree_test_batch: no communication, every rank writes the TSVs of its own pairs)."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)


def main():
    spec = json.load(open(sys.argv[1]))
    from scsommerclock_b200 import run_PT_test as P, sharding
    rank, world, _ = sharding.init_distributed()
    assert world == spec["world"]
    one = spec["one"]
    P.run_poisson_tree_test(one["vcf"], one["tree"], one["out"], spec["w_maxs"], "", "", None, None)
    sharding.barrier()
    sw = spec["sweep"]
    mine = P.run_poisson_tree_test_batch(sw["vcfs"], sw["trees"], sw["outs"], spec["w_maxs"])
    with open(spec["log"] + ".%d" % rank, "w") as f:
        json.dump({"rank": rank, "pairs": [m[0] for m in mine]}, f)
    sharding.barrier()
    import torch.distributed as dist
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
