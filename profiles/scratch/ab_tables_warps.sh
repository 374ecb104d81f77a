# one-off: tables_kernel with 8 / 12 / 24 warps per block (variant libraries built by hand)
for w in 8 24 28; do
  f=tdvmc_b200/libtdvmc_b200.so; [ $w != 8 ] && f=tdvmc_b200/libtdvmc_b200_w$w.so
  echo "warps=$w"; python - <<PY
import sys; sys.path.insert(0,'.')
from tdvmc_b200 import capi
capi.LIB_PATH = "$f"
exec(open('profiles/ab_tables.py').read())
PY
done
