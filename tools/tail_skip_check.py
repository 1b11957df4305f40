"""Development aid: acceptance sums of 1e7 thrown events with and without the walk's tail skip
(the library variant gpurun_variants/notail.so is built with NSS_NO_TAIL_SKIP) on several
configurations: passing-event counts must agree and the integrals to ~1e-7."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CASES = dict(c1={}, e10=dict(loge=10.0), e7=dict(loge=7.0), d33=dict(alt=33.0), w10=dict(ch=10.0),
             cloud3=dict(cloud=3.0))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import math
    import nuspacesim_b200 as nss
    from nuspacesim_b200.config import MonoCloud
    out = {}
    for name, kw in CASES.items():
        c = nss.NssConfig()
        c.simulation.spectrum.log_nu_energy = kw.get("loge", 8.0)
        c.detector.initial_position.altitude = kw.get("alt", 525.0)
        if "ch" in kw:
            c.simulation.max_cherenkov_angle = math.radians(kw["ch"])
        if "cloud" in kw:
            c.simulation.cloud_model = MonoCloud(altitude=kw["cloud"])
        r = nss.AcceptanceRunner(c, batch=1 << 20)
        r.enqueue(99, 0, 10_000_000)
        s = r.sums(False)
        out[name] = [float(x) for x in s[:6]]
    print("RESULT " + json.dumps(out))
else:
    res = {}
    var = os.environ.get("VAR", os.path.join(ROOT, "gpurun_variants"))
    for tag, lib in (("skip", None), ("noskip", os.path.join(var, "notail.so"))):
        env = dict(os.environ)
        if lib:
            env["NSS_B200_LIB"] = lib
        o = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        line = [l for l in o.stdout.splitlines() if l.startswith("RESULT ")]
        assert line, o.stderr[-2000:]
        res[tag] = json.loads(line[0][7:])
    for name in CASES:
        a, b = res["skip"][name], res["noskip"][name]
        print(f"{name:7s} n_pass {int(a[3])} vs {int(b[3])}  sum f rel diff {abs(a[1]-b[1])/max(abs(b[1]),1e-300):.2e}"
              f"  sum geo rel diff {abs(a[0]-b[0])/max(abs(b[0]),1e-300):.2e}  n_valid {int(a[4])}")
