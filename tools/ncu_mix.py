"""Instruction mix, stall reasons and barrier-delimited segments from an `ncu --page source --csv` dump."""
import csv, collections, sys

def load(path):
    rows = list(csv.reader(open(path)))
    kern = []; cur = None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}; kern.append(cur)
        elif cur is not None and cur["hdr"] is None: cur["hdr"] = r
        elif cur is not None and r: cur["rows"].append(r)
    # ncu prints each kernel twice (SASS view per source mode); keep the first of each name
    seen = set(); out = []
    for k in kern:
        if k["name"] in seen: continue
        seen.add(k["name"]); out.append(k)
    return out

def cls(o):
    if o in ("DFMA", "DADD", "DMUL"): return "fp64"
    if o in ("LDS", "STS", "LDG", "STG", "LDTM", "STTM", "LDL", "STL", "LDC", "LDCU", "LDSM"): return "mem"
    return "int"

def main(path, ops=None):
    for k in load(path):
        h = k["hdr"]; iS = h.index("Source"); iI = h.index("Instructions Executed"); iSm = h.index("# Samples")
        stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        tot = collections.Counter(); samp = collections.Counter(); st = collections.Counter()
        for r in k["rows"]:
            try: n = int(r[iI]); s = int(r[iSm])
            except ValueError: continue
            o = [t for t in r[iS].split() if not t.startswith("@")]
            opc = o[0].split(".")[0] if o else "?"
            tot[opc] += n; samp[opc] += s
            for i in stall:
                try: st[h[i]] += int(r[i])
                except ValueError: pass
        T = sum(tot.values()); S = sum(samp.values())
        print(k["name"][:70], "warp-instr", T, "samples", S)
        bycls = collections.Counter()
        for o, n in tot.items(): bycls[cls(o)] += n
        print("   classes:", {c: f"{n / T * 100:.1f}%" for c, n in bycls.items()})
        for o, n in tot.most_common(18): print(f"   {o:10s} {n / T * 100:5.1f}%  samples {samp[o] / S * 100:5.1f}%")
        ss = sum(st.values())
        print("   stalls:", ", ".join(f"{a[6:]} {b / ss * 100:.0f}%" for a, b in st.most_common(9)))

if __name__ == "__main__":
    main(sys.argv[1])


def segments(path, pattern, ops):
    """warp instructions per op between barriers / backward branches of the first kernel matching `pattern`"""
    k = [k for k in load(path) if pattern in k["name"]][0]
    h = k["hdr"]; iS = h.index("Source"); iI = h.index("Instructions Executed"); iSm = h.index("# Samples")
    S = 0
    for r in k["rows"]:
        try: S += int(r[iSm])
        except ValueError: pass
    acc = collections.Counter(); n_acc = 0; s_acc = 0; ops_c = collections.Counter()
    for idx, r in enumerate(k["rows"]):
        try: n = int(r[iI]); s = int(r[iSm])
        except ValueError: continue
        toks = [t for t in r[iS].split() if not t.startswith("@")]
        opc = toks[0].split(".")[0]
        acc[cls(opc)] += n; n_acc += n; s_acc += s; ops_c[opc] += n
        if opc == "BAR" or (opc == "BRA" and n_acc > ops * 150):
            top = ", ".join(f"{o} {c / ops:.0f}" for o, c in ops_c.most_common(6))
            print(f"{idx:5d} {toks[0][:10]:10s} instr/op {n_acc / ops:7.0f} fp64 {acc['fp64'] / ops:6.0f} int {acc['int'] / ops:6.0f} mem {acc['mem'] / ops:5.0f}"
                  f" samples {s_acc / S * 100:5.1f}% x{n / ops:.1f} | {top}")
            acc = collections.Counter(); n_acc = 0; s_acc = 0; ops_c = collections.Counter()
