for sm in 128 160 200 256; do
  echo "sparse_max=$sm"; python scripts/profile_one.py --rows 1000000000 --reps 2 --sparse-max $sm --thresholds 949,929,899,849 | tail -4
done
