# does a smaller ACS footprint (threads per SM, ring depth) let the other captures' kernels share the SMs?
for cfg in "384 6" "320 6" "256 0" "256 6" "256 4"; do
set -- $cfg
echo "tpsm $1 depth $2"
BENCH_QUICK=1 DVBT_B200_VIT_TPSM=$1 DVBT_B200_VIT_DEPTH=$2 timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | grep "bench quick"
done
