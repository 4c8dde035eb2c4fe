# A/B the product library against every libntm_b200_*.so (tools/ab_build.py) on the lean kernel's widths + accuracy vs the fp32 kernel
P=$PWD/neural-tape-modeling_b200
for l in libntm_b200.so $(cd $P; ls libntm_b200_*.so 2>/dev/null); do
  sfx=${l#libntm_b200}; sfx=${sfx%.so}
  echo "== lib$sfx"
  NTM_B200_LIB=$P/$l NTM_B200_TORCH_LIB=$P/ntm_b200_torch$sfx.so python tools/lean_check.py 2>&1 | grep "^f16 B="
  NTM_B200_LIB=$P/$l NTM_B200_TORCH_LIB=$P/ntm_b200_torch$sfx.so python tools/experiments/mma2_check.py f16 2>&1 | grep "B=  592\|B= 1024" | cut -c1-200
done
