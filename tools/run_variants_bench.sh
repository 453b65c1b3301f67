cd /root/repo
for f in variants/libspamm_*.so; do n=$(basename $f .so); echo "== $n"; SPAMM_B200_LIB=$PWD/$f python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-exact 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['ms_per_step_minmax'])"; done
