#!/bin/bash
# what does the GPU box hold?  (+ compute-sanitizer passes over the tiny fused fwd+bwd of smoke())
# usage (through gpurun): bash tools/gpu_probe.sh   -> gpurun_out/probe_*.txt
mkdir -p gpurun_out
{
  echo "== nvidia-smi"; nvidia-smi
  echo "== topo"; nvidia-smi topo -m
  echo "== nproc"; nproc; free -g | head -2
  echo "== /root/reference"; ls -la /root/reference 2>&1 | head; ls -laR /root/reference/submodules 2>&1 | head -20
  echo "== baseline/_ref, MEASURED_PEAKS.json"; ls -la baseline/_ref MEASURED_PEAKS.json 2>&1
  echo "== import diff_gaussian_rasterization / simple_knn"
  python -c "import diff_gaussian_rasterization as d; print('FOUND', d.__file__)" 2>&1 | tail -1
  python -c "import simple_knn; print('FOUND', simple_knn.__file__)" 2>&1 | tail -1
  echo "== glm.hpp"; find / -name glm.hpp -not -path '/proc/*' 2>/dev/null | head
  echo "== rasterizer sources anywhere"; find / \( -name 'rasterizer_impl*' -o -name 'diff_gaussian*' -o -name 'simple_knn*' \) -not -path '/proc/*' 2>/dev/null | head
} > gpurun_out/probe_step0.txt 2>&1
if [ "$1" != "--no-sanitizer" ]; then
  for tool in memcheck racecheck; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/probe_sanitizer_$tool.txt 2>&1
    echo "rc=$?" >> gpurun_out/probe_sanitizer_$tool.txt
  done
fi
tail -n 5 gpurun_out/probe_step0.txt; for f in gpurun_out/probe_sanitizer_*.txt; do tail -n 4 $f; done
