python bench.py > gpurun_out/r1_bench_n1.json 2> gpurun_out/r1_bench_n1.err; tail -1 gpurun_out/r1_bench_n1.json | python tools/benchline.py
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_ref.json 2>/dev/null; cut -c1-330 gpurun_out/r1_bench_ref.json
export ATSC_ENGINES=1 ATSC_WAVE_MI=128
CMD="python bench.py --steps 1 --warmup 3 --series 96 --no-cpu"
$CMD > gpurun_out/plain19.log 2>&1 && {
ncu --metrics gpu__time_duration.sum --clock-control none -s 31 -c 10 --csv --log-file gpurun_out/r1_final_launches.csv $CMD > /dev/null 2>&1
for k in k_poly k_fft_fwd k_stats; do
ncu --set full --clock-control none --import-source on -k $k -s 3 -c 1 -f -o gpurun_out/prof_r1_final_$k $CMD > gpurun_out/ncu19_$k.log 2>&1
done
}
grep -o "k_[a-z_0-9]*(.*ns\",\"[0-9]*" gpurun_out/r1_final_launches.csv | sed "s/(.*,\"/ /"
