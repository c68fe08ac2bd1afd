#!/bin/bash
# Build libptg_b200 with -DPTG_DEBUG_BOUNDS (every table index of the hot path checked on the device, violations
# reported through ptg_poll_error) and run the GPU suite + smoke against it.  Stand-in for compute-sanitizer, which
# is closed on this GPU pool.   usage (on a GPU box):  bash tools/debug_bounds.sh
set -e
cd "$(dirname "$0")/.."
if [ ! -f variants/dbg.so ]; then tools/build_variants.sh dbg:"-DPTG_DEBUG_BOUNDS"; fi
export PTG_B200_SO=$PWD/variants/dbg.so
python -m pytest tests -m gpu -x -q -k "not 1m_envs and not full_size and not whole_episode" 2>&1 | tail -5
python __graft_entry__.py smoke | tail -1
