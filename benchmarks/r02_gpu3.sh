for v in 1 8; do for c in 3 1; do echo "V=$v"; CSHIFT=$c ./benchmarks/micro/gen_harness_v$v 40000 | tail -1; done; done
