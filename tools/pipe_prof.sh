set -u
out=gpurun_out/r1d; mkdir -p $out
bash tools/pipeline_bench.sh $out 1000 5000 5000 > $out/pipeline.json 2> $out/pipeline.err
cat $out/pipeline.json; cat $out/pipe_summary.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/launches_pipeline1000.csv \
    nanocall_b200/bin/nanocall-b200 --pore r73 -o /tmp/p.fa --log warning /tmp/pipe.ncev > $out/lp.log 2>&1
python tools/make_synth_ncev.py /tmp/pipe_small.ncev 256 5000 5000 7 > /dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:st_stats_kernel -s 2 -c 1 -f -o $out/st_stats \
    nanocall_b200/bin/nanocall-b200 --pore r73 -o /tmp/p.fa --log warning /tmp/pipe_small.ncev > $out/ncu_st.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pm_stats_kernel -s 2 -c 1 -f -o $out/pm_stats \
    nanocall_b200/bin/nanocall-b200 --pore r73 -o /tmp/p.fa --log warning /tmp/pipe_small.ncev > $out/ncu_pm.log 2>&1
rm -f $out/pipe.fa $out/pipe.err
