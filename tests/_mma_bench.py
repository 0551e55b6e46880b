import sys; sys.path.insert(0, '.')
from replay_cql_b200.engine import CqlEngine, CqlHyperParams
eng = CqlEngine(CqlHyperParams(batch_size=64))
names = {0: "cg1 SS N128", 1: "cg1 TS N128", 2: "cg1 SS N256", 3: "cg1 TS N256", 4: "cg2 SS N128", 5: "cg2 TS N128", 6: "cg2 SS N256", 7: "cg2 TS N256"}
for three in (0, 8):
    for mode in range(8):
        for iters in (96, 768):
            eng.mma_bench(mode | three, iters)
            a, b = eng.mma_bench(mode | three, iters)
            print(f"{names[mode]:12s} {'3-term' if three else 'same  '} iters {iters:4d}: issue {a / iters:7.1f} clk/MMA   retire {b / iters:7.1f} clk/MMA")
eng.close()
