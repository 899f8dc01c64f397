"""Replays, in exact rational arithmetic, the two arithmetic shortcuts of common.cuh:
div_1e5_int53 (one FMA refinement step == the IEEE quotient n/1e5 for integer |n| < 2^53) and
round_half_away (trunc(y + copysign(pred(0.5), y)) == f64::round)."""
from fractions import Fraction as F
import random, math
B=100000.0; Y=1e-5
def fma(a,b,c): return float(F(a)*F(b)+F(c))
def one(n):
    q=n*Y
    r=fma(-B,q,n)
    return fma(r,Y,q)
random.seed(1)
bad=0;tot=0
def check(n):
    global bad,tot
    n=float(n); tot+=1
    if one(n)!=n/B:
        bad+=1; print("BAD",n)
for bits in range(1,54):
    for _ in range(3000):
        n=random.getrandbits(bits)|(1<<(bits-1))
        check(n); check(-n)
# near multiples of 3125 and near midpoints
for _ in range(40000):
    k=random.getrandbits(random.randint(1,40)); check(k*3125+random.randint(-2,2))
print(tot,bad)
# rounding trick
P=0.49999999999999994
def rnd(y): return float(math.trunc(y+math.copysign(P,y))) if abs(y)<2**62 else y
def ref(y):
    r=float(math.trunc(y)); d=y-r
    if abs(d)>=0.5: r+=math.copysign(1.0,y)
    return r
import struct
bad=0
for _ in range(400000):
    e=random.randint(-5,54); m=random.random()+1; y=math.ldexp(m,e)*random.choice([-1,1])
    if random.random()<0.5:
        y=float(round(y))+random.choice([0.5,-0.5,0.49999999999999994,-0.49999999999999994, math.nextafter(0.5,1),0.25])
    if rnd(y)!=ref(y): bad+=1; print("BADR",y,rnd(y),ref(y))
print("round bad",bad)
