cd $GRAFT_REPO_ROOT
for cfg in "0 0" "1 0" "1 500" "1 1000" "1 1400" "1 2000" "0 1000"; do
  set -- $cfg
  DGP_B200_WARPMAP=$1 DGP_B200_SKEW=$2 python bench.py --steps 6 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['roofline']['categories_ms_per_step']
print('warpmap $1 skew $2: step %.2f fwd %.2f bwd %.2f' % (d['ms_per_step'], c['fused_fwd'], c['fused_bwd']))"
done
