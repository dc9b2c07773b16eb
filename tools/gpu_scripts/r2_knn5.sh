#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for cm in 1 2 4 6 10; do
for b in encoded random; do
RF_KNN_COOP_MAX=$cm timeout 600 python bench.py --workload retrieval --bank $b --no-cpu-baseline --steps 3 > /tmp/b.json 2>/dev/null
python -c "
import json
l=json.load(open('/tmp/b.json')); print('coop_max=$cm $b', round(l['value']), l['breakdown_ms']['knn'])"
done
done
