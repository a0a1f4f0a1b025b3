for rows in 262144 1048576 2097152 4194304 8388608 16777216; do
  for plan in fused two_pass; do
    echo "rows=$rows plan=$plan: $(python scripts/profile_one.py --rows $rows --plan $plan --reps 6 2>&1 | tail -4 | awk '{s=0; for(i=5;i<=NF;i++){gsub(/[\[\],]/,"",$i); s+=$i}; printf "%s:%.3f ", $1, s}')"
  done
done
