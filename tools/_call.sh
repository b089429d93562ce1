timeout 300 python tools/timeline_graph.py 2>&1 | grep -v Warn | head -45
