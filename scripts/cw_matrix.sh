for args in "512,512 80 4" "512,512 84 4" "512,512 79 4" "512,512 96 4" "512,512 70 4" "512,512 100 4" "512,512 150 4"; do
  echo "== $args"; timeout 300 python scripts/cw_debug.py $args 2>&1 | grep -E "pol_b2|particles"
done
