python -m pytest tests -q -m gpu 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --cg 200 > gpurun_out/r01_final_c2.json 2> gpurun_out/r01_final_c2.err; cut -c1-300 gpurun_out/r01_final_c2.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 30 --csv --log-file gpurun_out/r01_launches_final_hexp2poisson_n128.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:assemble_gram_mma -s 1 -c 1 -o gpurun_out/r01_final_mma_hexp2poisson python bench.py --grid 64 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:assemble_gram_team -s 1 -c 1 -o gpurun_out/r01_final_team_hexp2elast python bench.py --phys elasticity --grid 48 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:assemble_gram_team -s 1 -c 1 -o gpurun_out/r01_final_team_hexp4poisson python bench.py --p 4 --grid 32 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:assemble_gram_team -s 1 -c 1 -o gpurun_out/r01_final_team_tetp2elast python bench.py --phys elasticity --topo tet --grid 48 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
