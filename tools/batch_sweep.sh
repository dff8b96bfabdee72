# config 4 against the wavefront batch size (DRT_BATCH_SLOTS): the whole render and one rank's share of an 8-way sharded one
for s in 4194304 8388608 16777216 33554432; do
  echo "=== DRT_BATCH_SLOTS=$s"
  DRT_BATCH_SLOTS=$s python tools/shard_overhead.py 2>&1 | tail -4
done
