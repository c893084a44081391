#!/bin/bash
# Recipe for oracle/_ref/: the UNMODIFIED reference package, placed where it can travel to the GPU
# box (oracle/_ref/ is git-ignored, not gpurun-ignored), so that `bench.py --impl reference` and
# bench.py's cpu_baseline leg time the reference's own PyTorch CPU path
# (exp_mnist_resnet/save_kernel.py:21-24 -> cnn_gp/kernels.py:18-57) rather than the C port.
#
# The reference is pure Python (setup.py:14 has ext_modules=[]): "building" it is copying the two
# packages the path imports, byte for byte, from where they lie under /root/reference.  Nothing is
# edited; the numpy alias the reference still uses (cnn_gp/data.py:12 `np.int`) is shimmed by the
# caller (oracle/ref_cpu.py), not patched here.  Outputs go only into oracle/_ref/.
set -e
here=$(cd "$(dirname "$0")" && pwd)
src=${1:-/root/reference}
if [ ! -d "$src/cnn_gp" ]; then
  echo "make_ref: $src/cnn_gp not found (no reference on this box); keeping $here/_ref as it is" >&2
  exit 0
fi
rm -rf "$here/_ref"
mkdir -p "$here/_ref"
cp -r "$src/cnn_gp" "$src/configs" "$here/_ref/"
find "$here/_ref" -name __pycache__ -type d -exec rm -rf {} +
( cd "$src" && sha256sum cnn_gp/*.py configs/*.py ) > "$here/_ref/SHA256SUMS"
echo "make_ref: $(wc -l < "$here/_ref/SHA256SUMS") reference files -> $here/_ref"
