import sys, numpy as np
a, b = np.load(sys.argv[1]), np.load(sys.argv[2])
for k in ["sst", "heatflux", "global_tas", "CO2_concentration"]:
    x, y = a[k], b[k]
    ne = (x != y) & ~(np.isnan(x) & np.isnan(y))
    if not ne.any():
        print(k, "identical"); continue
    yrs = np.argmax(ne, axis=1)
    has = ne.any(axis=1)
    print(k, "members differing", int(has.sum()), "of", len(has), "first-diff year idx: min", int(yrs[has].min()), "median", int(np.median(yrs[has])))
    m = int(np.argmax(has)); j = int(yrs[m])
    print("  member", m, "year idx", j, repr(x[m, j]), repr(y[m, j]), "ulps ~", abs(x[m,j]-y[m,j])/np.spacing(abs(x[m,j])))
