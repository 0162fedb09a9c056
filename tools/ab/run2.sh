mkdir -p gpurun_out
for v in old g4; do
  JUSTPIC_LIB=$PWD/tools/ab/libs/$v.so timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:'k_move_gather|k_move_scatter' --launch-skip 4 --launch-count 2 --csv --log-file gpurun_out/nanvac_$v.csv python tools/time_phases.py --cells 256 --steps 3 --classify 1 > /dev/null 2>&1
  echo "== $v"; grep -v "^==" gpurun_out/nanvac_$v.csv | cut -d, -f5,13- | tail -8
done
