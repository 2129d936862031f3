python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py --no-cpu 2>&1 | tail -1 | cut -c1-260
