# A/B every libntm_b200_*.so next to the package (tools/ab_build.py) on the latency-regime widths
P=$PWD/neural-tape-modeling_b200
for l in $(cd $P; ls libntm_b200_*.so); do
  sfx=${l#libntm_b200}; sfx=${sfx%.so}
  echo "== $sfx"
  NTM_B200_LIB=$P/$l NTM_B200_TORCH_LIB=$P/ntm_b200_torch$sfx.so python tools/lone_check.py 2>&1 | grep "^f16 \|cfg3 DiffDelGRU 256 x 30 s f16:"
done
