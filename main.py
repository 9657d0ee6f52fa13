"""``python main.py`` — the reference's batch driver (main.py:1-207) on the B200 engine: reads config.json, runs the
12-level x epoch experiment queue as one sharded GPU job and writes results/data/<timestamp>/{config.json,
results.csv, summary.json}.  Multi-GPU: ``torchrun --nproc-per-node N main.py``."""
import uvs_b200

if __name__ == "__main__":
    uvs_b200.sweep.main()
