"""tools/make_fsm_kats.py -- writes tests/golden/fsm_kats.json.

The EXPECTED values below were traced BY HAND through KmerGutsJava.gatherHits / processSetOfHits (KGJ:457-514,
385-455); this script only serialises them.  Nothing here calls an oracle or the product: the file pins both.
hit = [pos, fI, oI, wt, avgOffFromEnd]; call = [start, end, count, fI, weighted]; otu = [[count, oI], ...].
"""
import json, os
import numpy as np

def H(pos, fI, oI=1, wt=1.0, avg=0):
    return [pos, fI, oI, wt, avg]

D = dict(min_hits=5, max_gap=200, min_weighted_hits=0, order_constraint=0)
kats = []
def kat(name, hits, calls, otu, note, **p):
    kats.append(dict(name=name, note=note, params={**D, **p}, hits=hits, calls=calls, otu=otu))

kat("A_basic", [H(p, 7, 3, 0.5) for p in (0, 10, 20, 30, 40)], [[0, 47, 5, 7, 2.5]], [[5, 3]],
    "five hits of one function -> one CALL; end = last pos + 7 (KGJ:400)")
kat("B_too_few", [H(p, 7) for p in range(4)], [], [], "4 < minHits at the end (KGJ:511)")
kat("C_two_runs", [H(p, 7, 1) for p in range(5)] + [H(p, 7, 2) for p in range(300, 305)],
    [[0, 11, 5, 7, 5.0], [300, 311, 5, 7, 5.0]], [[5, 2], [5, 1]],
    "gap > maxGap splits (KGJ:477-480); a later equal count bubbles ahead (<=, KGJ:432)")
kat("D_gap_equal_joins", [H(p, 7) for p in range(5)] + [H(204, 7), H(205, 7)], [[0, 212, 7, 7, 7.0]], [[7, 1]],
    "last+maxGap < pos is strict: a gap of exactly 200 joins (KGJ:478)")
kat("E_gap_plus_one_splits", [H(p, 7) for p in range(5)] + [H(205, 7), H(206, 7)], [[0, 11, 5, 7, 5.0]], [[5, 1]],
    "gap 201 splits; trailing 2 hits are dropped")
kat("F_singleton_foreign", [H(0, 7), H(1, 7), H(2, 9), H(3, 7), H(4, 7), H(5, 7)], [[0, 12, 5, 7, 5.0]], [[5, 1]],
    "a single foreign fI inside a run is tolerated and not counted (KGJ:391, 503-504)")
kat("G_pair_switch", [H(p, 7) for p in range(5)] + [H(p, 9) for p in range(5, 10)],
    [[0, 11, 5, 7, 5.0], [5, 16, 5, 9, 5.0]], [[10, 1]],
    "second 9 triggers processSetOfHits; the retained pair seeds the next run (KGJ:442-449)")
kat("H_switch_discards_short", [H(p, 7) for p in range(4)] + [H(p, 9) for p in range(5, 10)],
    [[5, 16, 5, 9, 5.0]], [[5, 1]], "first run has 4 hits when the switch fires -> no call, pair retained")
kat("I_alternating", [H(p, 7 if p % 2 == 0 else 9) for p in range(12)], [[0, 17, 6, 7, 6.0]], [[6, 1]],
    "no two consecutive equal foreign fIs -> never switches; only fI 7 hits are counted; end = last 7 (pos 10)+7")
kat("J_otu_top5", [H(p, 7, o) for p, o in enumerate([1, 2, 2, 3, 3, 3, 4, 5, 6, 6])], [[0, 16, 10, 7, 10.0]],
    [[3, 3], [2, 6], [2, 2], [1, 5], [1, 4]], "sixth OTU overwrites the last buffer entry (KGJ:419-421), then bubbles")
kat("K_fp32_sum", [H(p, 7, 1, float(np.float32(0.1))) for p in range(10)],
    [[0, 16, 10, 7, float(sum([np.float32(0.1)] * 10, np.float32(0)))]], [[10, 1]],
    "ten fp32 additions of 0.1f in list order = 1.0000001 (prints 1.000000)")
kat("L_pair_dropped_by_gap", [H(p, 7) for p in range(5)] + [H(5, 9), H(6, 9)] + [H(p, 9) for p in (400, 401, 402)],
    [[0, 11, 5, 7, 5.0]], [[5, 1]], "retained pair (2 hits) < minHits at the gap -> cleared (KGJ:481-483)")
kat("M_weight_gate", [H(p, 7, 1, 0.5) for p in range(5)], [], [], "fICount ok but 2.5 < minWeightedHits=3 (KGJ:397)",
    min_weighted_hits=3)
kat("N_weight_gate_pass", [H(p, 7, 1, 0.75) for p in range(4)] , [[0, 10, 4, 7, 3.0]], [[4, 1]],
    "-m 4 -M 3: 4 x 0.75 = 3.0 >= 3", min_hits=4, min_weighted_hits=3)
kat("O_small_gap", [H(0, 7), H(50, 7), H(100, 7), H(151, 7), H(152, 7), H(153, 7)], [], [],
    "-m 3 -g 50: 0,50,100 join (gaps == 50); 151 is 51 away -> process: 3 hits -> wait see expected", min_hits=3, max_gap=50)
# O traced: at 151: 100+50 < 151 -> size 3 >= 3 -> CALL 0 107 3 7; then 151,152,153 -> end: size 3 -> CALL 151 160 3 7
kats[-1]["calls"] = [[0, 107, 3, 7, 3.0], [151, 160, 3, 7, 3.0]]
kats[-1]["otu"] = [[6, 1]]
kats[-1]["note"] = "-m 3 -g 50: gaps of exactly 50 join, 51 splits; both runs reach minHits=3"
kat("P_order_ok", [H(p, 7, 1, 1.0, 100 - p) for p in (0, 10, 20, 30, 40)], [[0, 47, 5, 7, 5.0]], [[5, 1]],
    "-O: offsets consistent (|dpos - davg| = 0 <= 20, KGJ:490-494)", order_constraint=1)
kat("Q_order_reject_offset", [H(0, 7, 1, 1.0, 100), H(10, 7, 1, 1.0, 90), H(20, 7, 1, 1.0, 20), H(30, 7, 1, 1.0, 70),
                               H(40, 7, 1, 1.0, 60), H(50, 7, 1, 1.0, 50)], [[0, 57, 5, 7, 5.0]], [[5, 1]],
    "-O: hit at 20 has |10 - 70| = 60 > 20 -> not appended; the rest chain off the hit at 10", order_constraint=1)
kat("R_order_reject_foreign", [H(0, 7), H(1, 7), H(2, 9), H(3, 9), H(4, 7), H(5, 7), H(6, 7)], [[0, 13, 5, 7, 5.0]],
    [[5, 1]], "-O: foreign fI hits are not appended, so no switch happens (avg offsets all 0: |dpos - 0| <= 20)",
    order_constraint=1)
kat("S_no_order_same_input", [H(0, 7), H(1, 7), H(2, 9), H(3, 9), H(4, 7), H(5, 7), H(6, 7)], [], [],
    "same hits without -O: 9,9 switches (2 sevens: no call), then 7,7 switches back (2 nines: no call); 3 left")
kat("T_edge_20", [H(0, 7, 1, 1.0, 0), H(30, 7, 1, 1.0, -10), H(60, 7, 1, 1.0, 1), H(61, 7, 1, 1.0, 0), H(62, 7, 1, 1.0, -1),
                  H(63, 7, 1, 1.0, -2), H(64, 7, 1, 1.0, -3)], [[0, 71, 6, 7, 6.0]], [[6, 1]],
    "-O: |30 - 10| = 20 accepted; |30 - (-11)| = 41 rejected (hit 60); 61 vs 30: |31 - (-10)| = 41 rejected too ... see trace",
    order_constraint=1)
# T traced: list [0]; 30: d=(30-0)-(0-(-10))=20 -> ok [0,30]; 60: (30)-(-10-1)=41 -> rej; 61: (31)-(-10-0)=41 -> rej;
# 62: (32)-(-10+1)=41 rej; 63: 33-(-10+2)=41 rej; 64: 34-(-10+3)=41 rej -> list [0,30] -> 2 < 5 -> no call
kats[-1]["calls"] = []
kats[-1]["otu"] = []
kats[-1]["note"] = "-O: |30-10| = 20 is accepted (<=); every later hit is 41 off the last ACCEPTED hit -> rejected; 2 hits, no call"
kat("U_three_functions", [H(p, 7) for p in range(5)] + [H(p, 9) for p in range(5, 10)] + [H(p, 11, 2) for p in range(10, 16)],
    [[0, 11, 5, 7, 5.0], [5, 16, 5, 9, 5.0], [10, 22, 6, 11, 6.0]], [[10, 1], [6, 2]],
    "two successive pair switches")
json.dump(kats, open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "fsm_kats.json"), "w"), indent=1)
print(len(kats), "KATs written")
