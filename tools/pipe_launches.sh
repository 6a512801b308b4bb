out=gpurun_out/r1f; mkdir -p $out
python tools/make_synth_ncev.py /tmp/pipe.ncev 1000 5000 5000 5 > /dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/launches_pipeline1000.csv \
    nanocall_b200/bin/nanocall-b200 --pore r73 -o /tmp/p.fa --log warning /tmp/pipe.ncev > $out/lp.log 2>&1
