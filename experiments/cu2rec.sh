#!/usr/bin/env bash
# The reference's experiments/cu2rec.sh grid (3 data sets x 5 iteration counts x 2 factor counts)
# through bin/mf; results/<date>-<commit>.{txt,jsonl,md}.
cd "$(dirname "$0")/.." && python -c "import __graft_entry__ as g; g.build()" && python experiments/run_grid.py "$@"
