# ncu --set full of the learner's chain / heads kernels (one B200): bash tests/learner_ncu.sh r02z [kernels...]
R=${1:-r02z}; shift
mkdir -p gpurun_out
bash tests/learner_launches.sh ${R}_pre > /dev/null 2>&1   # writes /tmp/ll.py
for k in ${@:-chain_fwd chain_bwd heads_bwd}; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 5 -c 1 -f -o gpurun_out/${R}_$k python /tmp/ll.py > /dev/null 2>&1
done
ls -la gpurun_out/${R}_*.ncu-rep
