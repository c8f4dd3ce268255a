#!/usr/bin/env python
"""Reads the TRACE lines tools/att_lab prints (-DATT_TRACE build) and prints, per traced CTA, the per-tile
timeline of each role relative to the CTA's first event, plus mean intervals between tags in steady state."""
import sys, collections
ev = collections.defaultdict(list)
sm = {}
for line in open(sys.argv[1]):
    if not line.startswith("TRACE"): continue
    t = line.split()
    slot, s, role, tag, j, clk = int(t[2]), int(t[4]), int(t[6]), int(t[8]), int(t[10]), int(t[12])
    ev[slot].append((clk, role, tag, j)); sm[slot] = s
show = int(sys.argv[2]) if len(sys.argv) > 2 else 0
by_sm = collections.defaultdict(list)
for slot in ev: by_sm[sm[slot]].append(slot)
print("co-resident slots by SM:", {k: v for k, v in by_sm.items() if len(v) > 1})
for slot in sorted(ev):
    e = sorted(ev[slot], key=lambda x: (x[0] - ev[slot][0][0]) & 0xffffffff)
    t0 = min(x[0] for x in e)
    if slot == show:
        for clk, role, tag, j in sorted(e, key=lambda x: (x[0]-t0) & 0xffffffff):
            print(f"  slot {slot} +{(clk-t0)&0xffffffff:7d} role {role} tag {tag:2d} j {j}")
    # steady-state period: softmax tag 21 (S ready) between consecutive j
    s21 = {j: clk for clk, role, tag, j in e if tag == 21}
    js = sorted(s21)
    if len(js) > 3:
        per = [(s21[js[i+1]] - s21[js[i]]) & 0xffffffff for i in range(1, len(js)-1)]
        print(f"slot {slot} sm {sm[slot]}: tiles {len(js)} period mean {sum(per)/len(per):.0f} min {min(per)} max {max(per)} total {(max(x[0] for x in e)-t0)&0xffffffff}")
# mean deltas between consecutive events per role in steady state (tiles 2..n-2)
def deltas(role_tags):
    acc = collections.defaultdict(list)
    for slot in ev:
        idx = {(tag, j): clk for clk, role, tag, j in ev[slot]}
        for (a, da), (b, db), name in role_tags:
            for j in range(2, 10):
                if (a, j+da) in idx and (b, j+db) in idx:
                    acc[name].append((idx[(b, j+db)] - idx[(a, j+da)]) & 0xffffffff)
    for name, v in acc.items():
        v.sort()
        print(f"  {name:46s} mean {sum(v)/len(v):7.0f}  p10 {v[len(v)//10]:6d} p50 {v[len(v)//2]:6d} p90 {v[9*len(v)//10]:6d}")
deltas([((20,0),(21,0),"softmax: wait s_full (20->21)"),
        ((21,0),(22,0),"softmax: S tmem ld (21->22)"),
        ((22,0),(23,0),"softmax: max + exps chunk0 (22->23)"),
        ((23,0),(24,0),"softmax: wait o_full (23->24)"),
        ((22,0),(25,0),"softmax: compute+P stores (22->25)"),
        ((25,0),(26,0),"softmax: st wait (25->26)"),
        ((21,0),(21,1),"softmax: tile period (21->21')"),
        ((26,0),(15,0),"P stored -> MMA warp past P_FULL (26->15)"),
        ((22,0),(11,1),"S drained -> MMA warp past S_EMPTY (22->11')"),
        ((12,1),(21,1),"QK issued -> softmax sees s_full (12'->21')"),
        ((16,0),(24,1),"PV issued -> softmax sees o_full next tile (16->24')"),
        ((11,0),(12,0),"MMA: issue QK (11->12)"),
        ((15,0),(16,0),"MMA: issue PV (15->16)"),
        ((13,0),(11,0),"MMA: wait S_EMPTY (13->11)"),
        ((14,0),(15,0),"MMA: wait P_FULL (14->15)"),
        ((16,0),(13,2),"MMA: wait k_full next (16->13'')"),
        ((12,1),(14,0),"MMA: wait v_full (12'->14)")])
