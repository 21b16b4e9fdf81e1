for v in "" mb7 ""; do
  echo "== variant '$v'"
  if [ -n "$v" ]; then export NEXUS_B200_LIB=$PWD/nexus_b200/variants/lib_$v.so; else unset NEXUS_B200_LIB; fi
  NX_FRAMES=4 timeout 300 python scripts/tune_pool.py instanced10m_4k lane lane 2>&1 | grep -v "^      "
done
