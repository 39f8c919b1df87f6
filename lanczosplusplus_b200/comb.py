"""The `.comb` files of `lanczos -g` (LanczosDriver1.h:147-181): same layout as host/comb_io.h (see the note there on what is
fixed by the reference's in-repo consumers and what is this repository's choice)."""


def write_comb(path, site0, site1, index_to_cf, cfs):
    """cfs: ContinuedFraction objects (a, b, eg, weight, isign); index_to_cf: "spin,type,orb1,orb2" strings (Engine.h:199-202)."""
    with open(path, "w") as f:
        f.write("#Site0=%d\n#Site1=%d\n#INDEXTOCF %s \n" % (site0, site1, " ".join(index_to_cf)))
        f.write("#CONTINUEDFRACTIONCOLLECTION=%d\n" % len(cfs))
        for cf in cfs:
            f.write("#Avector\n%d\n" % len(cf.a) + "".join("%.17g\n" % v for v in cf.a))
            f.write("#Bvector\n%d\n" % len(cf.b) + "".join("%.17g\n" % v for v in cf.b))
            f.write("#CFWeight=%.17g\n#CFEnergy=%.17g\n#CFIsign=%d\n" % (cf.weight, cf.eg, cf.isign))


def read_comb(path):
    """-> dict(site0, site1, index_to_cf, cfs=[dict(a, b, weight, eg, isign)])"""
    out = dict(site0=0, site1=0, index_to_cf=[], cfs=[])
    lines = open(path).read().splitlines()
    i = 0

    def vector(i):
        n = int(lines[i])
        return [float(x) for x in lines[i + 1:i + 1 + n]], i + 1 + n

    while i < len(lines):
        ln = lines[i]
        i += 1
        if ln.startswith("#Site0="):
            out["site0"] = int(ln[7:])
        elif ln.startswith("#Site1="):
            out["site1"] = int(ln[7:])
        elif ln.startswith("#INDEXTOCF"):
            out["index_to_cf"] = ln[10:].split()
        elif ln.startswith("#Avector"):
            a, i = vector(i)
            out["cfs"].append(dict(a=a, b=[], weight=0.0, eg=0.0, isign=1))
        elif ln.startswith("#Bvector"):
            out["cfs"][-1]["b"], i = vector(i)
        elif ln.startswith("#CFWeight="):
            out["cfs"][-1]["weight"] = float(ln[10:])
        elif ln.startswith("#CFEnergy="):
            out["cfs"][-1]["eg"] = float(ln[10:])
        elif ln.startswith("#CFIsign="):
            out["cfs"][-1]["isign"] = int(ln[9:])
    return out
