#!/usr/bin/env bash
# The reference's experiments/cu2rec_prof.sh (nvprof per run) with ncu: launch lists under results/prof/.
cd "$(dirname "$0")/.." && python -c "import __graft_entry__ as g; g.build()" && \
    python experiments/run_grid.py --prof --datasets ml-100k ml-20m "$@"
