#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 40 python -m pytest tests -m gpu -q -x -k "not full_size and not spiral_10k and not default_mode_trace" > gpurun_out/memcheck_full.log 2>&1
grep -E "^========= (Invalid|Program hit|Error|ERROR|Uninit|Misaligned|Out-of|Illegal|Leak|Barrier|Race)|^=========     (at |by thread|Address|and is)|Host Frame: (yh_|yref_|[a-z_]+ in (oracle_lib|host|io|test_))" gpurun_out/memcheck_full.log | head -80
tail -4 gpurun_out/memcheck_full.log
