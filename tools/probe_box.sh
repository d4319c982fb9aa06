#!/bin/bash
# One-off probe of the GPU box: is R / RANN / Seurat / Eigen / ANN present
# so that real nn2 + ComputeSNN goldens could be dumped?  Output: gpurun_out/r02_probe_box.txt
out=gpurun_out/r02_probe_box.txt
mkdir -p gpurun_out
{
  echo "== date"; date -u
  echo "== R toolchain"
  for b in Rscript R r littler; do printf '%s: ' $b; command -v $b || echo "not found"; done
  ls -d /usr/lib/R /usr/local/lib/R /usr/lib64/R /opt/R 2>&1
  echo "== libR / RANN / ANN / Eigen anywhere on the filesystem"
  find / -xdev \( -name 'libR.so*' -o -iname 'RANN*' -o -name 'libANN*' -o -name 'ANN.h' -o -path '*/Eigen/Sparse' -o -name 'RcppEigen*' -o -iname 'Seurat*' \) 2>/dev/null | grep -v "^/root/repo\|gpurun" | head -20
  echo "(end of find)"
  echo "== python bridges"
  python - <<'PY'
for m in ("rpy2", "anndata", "scanpy", "pynndescent", "annoy", "hnswlib", "faiss", "sklearn", "scipy"):
    try:
        __import__(m); print(m, "importable")
    except Exception as e:
        print(m, "absent:", type(e).__name__)
PY
  echo "== host"
  nproc; lscpu | grep -E "Model name|Socket|Thread|Core" ; free -g | head -2
  echo "== gpus"
  nvidia-smi --query-gpu=index,name,memory.total,clocks.max.sm --format=csv
  nvidia-smi topo -m 2>&1 | head -20
  echo "== ncu"
  command -v ncu; ncu --version 2>&1 | tail -1
} > $out 2>&1
cat $out
