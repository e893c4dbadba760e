#!/usr/bin/env python
"""What the step kernel's hot loop really executes, from SASS: FP64 instruction counts per physics step.
usage: python tools/hot_loop_flops.py [lib.so]   -> per NC: DFMA/DMUL/DADD/DSETP/MUFU counts and executed flops (FMA = 2)
Importable: count(lib) -> {nc: {"executed_flops": .., "dfma": .., ...}} (bench.py reports it as roofline.executed_flops)."""
import collections, os, re, subprocess, sys

DEFAULT_LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cdpr_simulation_b200", "libcdpr_b200.so")


def count(lib=DEFAULT_LIB):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    out = {}
    for nc in (8, 4):
        # the kernel bench.py's headline runs: velocity mode, moment D-term, every robot-constant specialisation of the default
        # robot with that many cables (SPEC 63 = ... | SPEC_PAIR for the 8-cable cube, 31 for the 4-cable reference robot)
        funcs = re.split(r"\n\s*Function : ", txt)[1:]
        cand = [x for spec in (63, 31) for x in funcs if f"k_step_fastILi{nc}ELi11ELi2ELb1ELi{spec}E" in x.split("\n", 1)[0]]
        f = cand[0]
        ins = [(int(m.group(1), 16), m.group(2).strip()) for m in re.finditer(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);", f)]
        loops = []
        for a, t in ins:
            m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a:
                loops.append((int(m.group(1), 16), a))
        best = None
        for lo, hi in loops:  # innermost loop with the most DFMA = the hot loop
            if any(lo <= l2 and h2 <= hi and (l2, h2) != (lo, hi) for l2, h2 in loops):
                continue
            # the rare saturated pass hangs off the warp vote: [@!P BRA target] right after VOTE.ANY skips it; leave it out
            cold = (0, 0)
            body = [(a, t) for a, t in ins if lo <= a <= hi]
            for j, (a, t) in enumerate(body):
                if t.startswith("VOTE.ANY"):
                    for a2, t2 in body[j + 1:j + 60]:
                        m2 = re.match(r"@!?P\d+\s+BRA\s+0x([0-9a-f]+)", t2)
                        if m2:
                            cold = (a2 + 1, int(m2.group(1), 16))
                            break
            c = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for a, t in body if not (cold[0] <= a < cold[1]))
            # the hot loop = the smallest big loop that calls the out-of-line saturated pass (the optimistic body without
            # the rollout cost); the body that clamps inline has no call
            if c["DFMA"] < 100 or not any("CALL" in t for a, t in body):
                continue
            if best is None or sum(c.values()) < best[2]:
                best = ((lo, hi), c, sum(c.values()))
        (lo, hi), c, nb = best
        fp = c["DFMA"] + c["DMUL"] + c["DADD"] + c["DSETP"]
        out[nc] = {"range": (lo, hi), "instructions": nb, "dfma": c["DFMA"], "dmul": c["DMUL"], "dadd": c["DADD"], "dsetp": c["DSETP"],
                   "mufu": c["MUFU"], "lds": c["LDS"], "sts": c["STS"], "ldcu": c["LDCU"], "fp64_pipe": fp,
                   "executed_flops": 2 * c["DFMA"] + c["DMUL"] + c["DADD"] + c["MUFU"], "issue_slots": 2 * fp + (nb - fp)}
    return out


if __name__ == "__main__":
    for nc, r in count(sys.argv[1] if len(sys.argv) > 1 else DEFAULT_LIB).items():
        print(f"NC={nc}: hot loop {r['range'][0]:#x}..{r['range'][1]:#x}, {r['instructions']} instructions; DFMA {r['dfma']} DMUL {r['dmul']} "
              f"DADD {r['dadd']} DSETP {r['dsetp']} MUFU {r['mufu']} -> {r['fp64_pipe']} FP64-pipe instructions, {r['executed_flops']} executed flops "
              f"(FMA = 2); LDS {r['lds']} STS {r['sts']} LDCU {r['ldcu']}; issue slots 2 x FP64 + other = {r['issue_slots']}")
