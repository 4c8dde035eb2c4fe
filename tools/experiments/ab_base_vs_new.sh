P=$PWD/neural-tape-modeling_b200
for sfx in "_base" ""; do
  echo "== lib$sfx"
  export NTM_B200_LIB=$P/libntm_b200$sfx.so NTM_B200_TORCH_LIB=$P/ntm_b200_torch$sfx.so
  python tools/lone_check.py 2>&1 | grep "^f16 \|^bf16 \|cfg3 DiffDelGRU 256 x 30 s f16:"
done
