# captures of ranks 4..7 of an 8-GPU run: any acquisition behaviour the tables do not cover?
for sd in 5 6 7 8; do
echo "seed $sd"
DVBT_B200_ACQ_TRACE=1 BENCH_QUICK=1 BENCH_VERBOSE=1 BENCH_SEED=$sd timeout 600 python bench.py --steps 6 --warmup 3 2>&1 | grep "bench quick\|stages:\|acq batch" | sort | uniq -c | sort -rn | head -4 | cut -c1-260
done
