# A/B of prebuilt library variants on ONE box: VARS="A B" WLS="s64 s150d3" bash tools/ab_variants.sh  (build/variants/lib_<X>.so are swapped in for libnwsolve.so)
cp nahrwhals_b200/csrc/libnwsolve.so /tmp/orig.so
for v in $VARS; do
  cp build/variants/lib_$v.so nahrwhals_b200/csrc/libnwsolve.so
  for w in $WLS; do
    timeout 300 python tools/profile_solve.py $w 5 > /tmp/p.log 2>&1
    echo "$v $w $(grep 'rep 4' /tmp/p.log | cut -c1-75)"
  done
done
cp /tmp/orig.so nahrwhals_b200/csrc/libnwsolve.so
