import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        m = d.get("mcmc", {})
        print(f, "value %.3e" % d["value"], "kernel_ms %.4f" % d["roofline"]["kernel_ms"], "frac %.3f" % d["roofline"]["frac"],
              "e2e %.3e" % d["e2e"]["value"], "mcmc %.3e (%.3f ms/it)" % (m.get("value", 0), m.get("ms_per_iter", 0)), d["clocks"])
    except Exception as ex:
        print(f, "ERR", ex, open(f).read()[-1500:])
