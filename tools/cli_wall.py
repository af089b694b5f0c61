import sys, json; sys.path.insert(0, '/root/repo')
import bench
ref, sites, gts, rs, regions, graphs, batches = bench.make_workload(0)
print(json.dumps(bench.cli_end_to_end(ref, sites, rs, regions)))
