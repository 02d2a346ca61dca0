"""Near-minimax polynomial of lp_acos_unit (csrc/lp_internal.cuh): acos(1 - t) = sqrt(2 t) (1 + t P(t)) on [0, 0.4375]."""
import mpmath as mp
mp.mp.dps = 60
# acos(1 - t) = sqrt(2 t) * (1 + t * P(t)),  t in [0, T]
T = mp.mpf('0.4375')
def P(t):
    t = mp.mpf(t)
    if t < mp.mpf('1e-25'):
        return mp.mpf(1)/12 + 3*t/160
    return (mp.acos(1 - t) / mp.sqrt(2*t) - 1) / t
for deg in (10, 11, 12):
    coeffs, err = mp.chebyfit(P, [0, T], deg + 1, error=True)
    print(deg, mp.nstr(err, 5))
coeffs, err = mp.chebyfit(P, [0, T], 12, error=True)   # degree 11, highest first
print([float(c).hex() for c in coeffs])
import struct
open(__import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), 'acos_coef.h'),'w').write("static const double ACOS_P[12] = {\n" + ",\n".join("    %s /* %.20e */" % (float(c).hex(), float(c)) for c in reversed(coeffs)) + "\n};\n")
