for flags in "" "-DMDSEG_BWD_XP_NOLD=1"; do
  touch mul-datasets-semantic-segmentation_b200/csrc/mds_bwd.cu
  MDSEG_CFLAGS="$flags" python mul-datasets-semantic-segmentation_b200/build.py > /dev/null
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mds_bwd --csv --log-file gpurun_out/l_tmp.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-kernel-times > /dev/null 2>&1
  echo "== $flags"; python profiles/summarize.py launches gpurun_out/l_tmp.csv /tmp/x.md; grep "unnamed" /tmp/x.md
done
