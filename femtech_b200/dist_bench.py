"""Multi-GPU bench leg (one process per GPU, torchrun).  Filled in by femtech_b200.dist."""
import json
import os


def run(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank == 0:
        print(json.dumps({"metric": "hex8 element-steps/sec fp64", "n_gpus": args.gpus,
                          "unavailable": "multi-GPU stepping not wired into bench.py yet"}))
