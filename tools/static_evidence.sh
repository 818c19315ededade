#!/bin/bash
# Static evidence for profiles/: ptxas -v (registers, spills, stack, shared memory) of every kernel of the built
# library sources, and SASS statistics (instruction count to the first EXIT, opcode mix) of the hot kernels.
#   tools/static_evidence.sh TAG      ->  profiles/TAG_ptxas.txt, profiles/TAG_sass_stats.md
set -e
tag=${1:-r2}
root=$(cd "$(dirname "$0")/.." && pwd)
cd "$root/gym_uav_collision_avoidance_b200/csrc"
out="$root/profiles/${tag}_ptxas.txt"
echo "# nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xptxas -v  (uavca_kernels.cu uavca_policy.cu)" > "$out"
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xptxas -v -c -o /dev/null uavca_kernels.cu 2>&1 \
  | grep -A2 "Compiling entry function" | grep -v "^--" | c++filt >> "$out" || true
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xptxas -v -c -o /dev/null uavca_policy.cu 2>&1 \
  | grep -A2 "Compiling entry function" | grep -v "^--" | c++filt >> "$out" || true
md="$root/profiles/${tag}_sass_stats.md"
echo "# SASS statistics of the hot kernels (tools/sass_stats.py on the built libuavca.so; static counts up to the first unconditional EXIT = the common path + its rare branches)" > "$md"
cd "$root"
for k in 'step_multi_kernelILi8E' 'step_multi_kernelILi32E' 'step_multi_ring_kernelILi10E' 'step_multi_cta_kernel' 'rollout_multi_cta_kernel' 'policy_act_kernel' 'rollout_multi_kernelILi8E' 'rollout_multi_kernelILi32E' 'step_single_kernel' 'rollout_single_kernel'; do
  echo -e "\n## $k\n\`\`\`" >> "$md"
  python tools/sass_stats.py gym_uav_collision_avoidance_b200/libuavca.so "$k" | head -2 >> "$md"
  echo '```' >> "$md"
done
echo "wrote $out $md"
