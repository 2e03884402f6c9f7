for cfg in "" "LR_CONV_EPI_GROUPS=2" "LR_CONV_EPI_GROUPS=1" "LR_CONV_SETS=1" "LR_CONV_ASETS=1" "LR_CONV_ISSUERS=3" "LR_CONV_SEAM=2"; do
  env $cfg python bench.py --no-kernels --no-cpu-baseline --steps 5 > gpurun_out/sw.json 2>/dev/null
  python -c "
import json; d=json.loads(open('gpurun_out/sw.json').read()); print('$cfg'.ljust(24), round(d['ms_per_step'],3), {k:round(v['ms'],3) for k,v in d['roofline']['per_launch'].items()})"
done
