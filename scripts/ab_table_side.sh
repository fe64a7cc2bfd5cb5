for t in 1 0 1 0; do echo "TABLE_SIDE=$t"; LINKB200_TABLE_SIDE=$t timeout 100 python scripts/step_graph_probe.py 2>&1 | tail -2; done
