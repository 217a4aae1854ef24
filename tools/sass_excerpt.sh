#!/bin/bash
# The bulk-copy (TMA unit) and mbarrier instructions of the partition kernels in the built library:
# UBLKCP = cp.async.bulk global -> shared, SYNCS.* = mbarrier init / arrive.expect_tx / try_wait.
LIB=${1:-real_b200/libreal_gpu.so}
for f in _ZN7realgpu11k_part_histENS_10ScanParamsE _ZN7realgpu14k_part_scatterILb0EEEvNS_10ScanParamsE; do
  echo "== $f"
  cuobjdump -sass -fun "$f" "$LIB" 2>/dev/null | grep -E "UBLKCP|SYNCS|Function"
done
echo "== instruction mix of k_bucket_probe<false,true> (count by mnemonic, top 25)"
cuobjdump -sass -fun _ZN7realgpu14k_bucket_probeILb0ELb1EEEvNS_10ScanParamsE "$LIB" 2>/dev/null | grep -oE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T] )?[A-Z0-9_.]+" | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -25
