set -x
python bench.py --steps 10 --warmup 3 --cg 50 > gpurun_out/r01_c2.json 2> gpurun_out/r01_c2.err; cut -c1-400 gpurun_out/r01_c2.json; tail -2 gpurun_out/r01_c2.err
python bench.py --steps 5 --warmup 3 --p 4 --grid 64 --no-e2e > gpurun_out/r01_c4.json 2> gpurun_out/r01_c4.err; cut -c1-400 gpurun_out/r01_c4.json; tail -2 gpurun_out/r01_c4.err
python bench.py --steps 5 --warmup 3 --phys elasticity --grid 80 --no-e2e > gpurun_out/r01_c5share.json 2> gpurun_out/r01_c5share.err; cut -c1-400 gpurun_out/r01_c5share.json; tail -2 gpurun_out/r01_c5share.err
python bench.py --steps 5 --warmup 3 --p 1 --grid 32 --cg 50 > gpurun_out/r01_c1.json 2> gpurun_out/r01_c1.err; cut -c1-400 gpurun_out/r01_c1.json; tail -2 gpurun_out/r01_c1.err
timeout 900 python bench.py --steps 5 --warmup 3 --phys elasticity --topo tet --grid 113 --no-e2e > gpurun_out/r01_c3.json 2> gpurun_out/r01_c3.err; cut -c1-400 gpurun_out/r01_c3.json; tail -2 gpurun_out/r01_c3.err
ncu --set full --clock-control none --import-source on -k regex:assemble_gram_team -s 1 -c 1 -o gpurun_out/r01_team_p4 python bench.py --p 4 --grid 32 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_p4.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file gpurun_out/r01_launches_hexp4poisson_n32.csv python bench.py --p 4 --grid 32 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:assemble_gram_team -s 1 -c 1 -o gpurun_out/r01_team_elast_p2 python bench.py --phys elasticity --grid 48 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_el.log 2>&1
