for g in 1000 8 12 16; do
GPEDM_DAG_LAUUM_GROUP=$g timeout 200 python tools/dag_sweep.py 4096 8192 16384 2>&1 | cut -c1-140
GPEDM_DAG_LAUUM_GROUP=$g ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:dag_factor -s 2 -c 1 python tools/one_eval.py 8192 3 2>&1 | grep -E "dram__|gpu__time|lts__"
done
