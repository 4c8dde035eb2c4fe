P=$PWD/neural-tape-modeling_b200
for l in $(cd $P; ls libntm_b200*.so); do
  sfx=${l#libntm_b200}; sfx=${sfx%.so}
  NTM_B200_LIB=$P/$l NTM_B200_TORCH_LIB=$P/ntm_b200_torch$sfx.so timeout 300 python tools/lean_sweep.py 2>&1 | tail -2
done
