#!/bin/bash
# bench_brief for the default library and every eagle-mpc_b200/lib/libvar_*.so (kernel-variant experiments)
echo "== default"; bash scripts/bench_brief.sh
for v in eagle-mpc_b200/lib/libvar_*.so; do echo "== $v"; EMPC_LIB=$PWD/$v bash scripts/bench_brief.sh; done
