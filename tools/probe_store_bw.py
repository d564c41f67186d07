import sys; import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0]=[R, R+'/tests']
import numpy as np, verkko_hem_repo_b200 as vh
from helpers import b_phase_state, coef_vector
m=vh.unit_cube(1,5,half=20.0); T=m.tables(0)
ctx=vh.Context(T); ctx.set_coef_vector(coef_vector()); ctx.set_solution(b_phase_state(T,noise=0.0)); ctx.assemble()
nb=8*324*ctx.info()['nnzb']
for what in (6,7,8,6):
    for fl in (True,False):
        ms=ctx.time_kernel(what,reps=10,flush_l2=fl)
        print("what",what,"flush",fl,"ms %.4f"%ms,"GB/s %.0f"%(nb/ms/1e6))
    if what in (7,8): ctx.assemble()
