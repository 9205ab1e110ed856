#!/bin/bash
# e2e chunk schedule A/B + sanitizer over every kernel family. usage: tools/gpu_round_b.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -3
for rep in 1 2; do for hc in 1 2; do
  echo -n "c4 HOST_CHUNKS=$hc: "
  POLAR_B200_HOST_CHUNKS=$hc python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 5 2>>gpurun_out/e2e_$tag.err | tee -a gpurun_out/e2e_$tag.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), d['e2e']['pipelined_chunks'], d['e2e']['matches_device_arm'])"
done; done
for c in c5 c3; do for hc in 1 2; do
  echo -n "$c HOST_CHUNKS=$hc: "
  POLAR_B200_HOST_CHUNKS=$hc python bench.py --config $c --steps 5 --warmup 3 --no-cpu --e2e-steps 5 2>>gpurun_out/e2e_$tag.err | tee -a gpurun_out/e2e_$tag.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), d['e2e']['pipelined_chunks'], d['e2e']['matches_device_arm'])"
done; done
timeout 400 compute-sanitizer --tool memcheck python tools/sanitize_probe.py 2>&1 | tail -12 > gpurun_out/sanitizer_$tag.txt; timeout 400 compute-sanitizer --tool racecheck python tools/sanitize_probe.py 2>&1 | tail -3 >> gpurun_out/sanitizer_$tag.txt; cat gpurun_out/sanitizer_$tag.txt
