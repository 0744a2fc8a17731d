#!/bin/bash
# flood distance-window sweep under the concurrent batch load (throughput, not latency): VF_FLOOD_LEVELS x bench.py --workload batch
for lv in "$@"; do
  VF_FLOOD_LEVELS=$lv timeout 100 python bench.py --workload batch --meshes 160 --warmup 2 --jobs 16 2>/dev/null | tail -1 > /tmp/lv.json
  python -c "import json; d=json.load(open('/tmp/lv.json')); print('levels', $lv, 'models/s', round(d['value'], 1), 'checksum', d['config']['checksum_rank0'])"
done
