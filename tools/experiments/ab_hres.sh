P=$PWD/neural-tape-modeling_b200
for sfx in "" "_hres"; do
  echo "== lib$sfx"
  export NTM_B200_LIB=$P/libntm_b200$sfx.so NTM_B200_TORCH_LIB=$P/ntm_b200_torch$sfx.so
  python tools/lone_check.py 2>&1 | grep "^f16 \|cfg3 DiffDelGRU 256 x 30 s f16:"
  rm -f gpurun_out/parity_10s.json
  python -m pytest tests/test_parity_10s_gpu.py -m gpu -q -k "f16-kernel" 2>&1 | tail -1
  python - <<'PY'
import json
for r in json.load(open('gpurun_out/parity_10s.json')):
    if r['mode']=='f16':
        print(r['config'], r['kernel'][:14], {k:(round(v,8) if isinstance(v,float) else [float('%.2g'%x) for x in v]) for k,v in r.items() if k.startswith('esr') or k.startswith('max_abs')})
PY
done
