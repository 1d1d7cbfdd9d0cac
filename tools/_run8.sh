for c in 8 16 24 32 48 64 96; do
  for w in s150 s64; do
    NW_DP_COOP_MAX_CTAS=$c timeout 300 python tools/profile_solve.py $w 5 > /tmp/p.log 2>&1
    echo "coop_max $c $w $(grep 'rep 4' /tmp/p.log | cut -c1-75)"
  done
done
