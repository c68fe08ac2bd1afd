#!/usr/bin/env bash
# Stage the UNMODIFIED reference (SimMarkt/RL_PtG, MIT) into the git-ignored baseline/_ref/ so that it travels to the
# GPU box with gpurun (which ships /root/repo only).  bench.py's cpu_baseline leg and `--impl reference` import
# env/ptg_gym_env.py from there through the gymnasium stub of oracle/ref_harness.py and time it on the box's host
# cores.  Nothing under baseline/_ref is product source and
# nothing there is ever committed (.gitignore); the product never imports it.
set -euo pipefail
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
DST="$HERE/baseline/_ref"
if [ ! -f "$SRC/env/ptg_gym_env.py" ]; then
    echo "stage_reference: no reference checkout at $SRC (nothing staged)" >&2
    exit 0
fi
mkdir -p "$DST"
for d in env src config; do
    rm -rf "$DST/$d"
    cp -r "$SRC/$d" "$DST/$d"
done
cp "$SRC/LICENSE" "$DST/LICENSE"
find "$DST" -name '__pycache__' -type d -prune -exec rm -rf {} +
( cd "$SRC" && sha256sum env/ptg_gym_env.py src/rl_utils.py src/rl_opt.py ) > "$DST/SHA256SUMS"
echo "staged $(du -sh "$DST" | cut -f1) of reference sources into $DST"
