import json, sys
for n in sys.argv[1:]:
    try:
        d=json.loads(open("gpurun_out/ab_%s.json"%n).read().strip().splitlines()[-1])
        st=d["roofline"]["stage_ms"]
        print(n, "step %.3f e2e %.3f osc %.3f ends %.3f noise %.3f | held %.3f"%(d["ms_per_step"], d["e2e"]["ms_per_step"], st["oscillators"], st["phase_ends"], st["noise"], d["held_notes_variant"]["ms_per_step"]))
    except Exception as e:
        print(n, "failed", e)
