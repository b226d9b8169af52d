import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-38s %8.1f Mpix/s %7.3f ms  e2e %7.1f | "%(f.split('/')[-1], d["value"], d["ms_per_step"], d["e2e"]["value"]) + "  ".join("%s %.3f"%(k[:14],v["ms_per_launch"]*v["launches_per_frame"]) for k,v in d["kernels"].items()))
    except Exception as e: print(f,"ERR",e)
