#!/bin/bash
# Round 2, call AR (1 GPU): the whole GPU suite and smoke on the last commit.
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
