#!/usr/bin/env python3
"""Generate the F-16 model data/wiring files from the reference's JSBSim XML.

Inputs (read-only, only available in the build container):
    <ref>/aircraft/f16/f16.xml                  metrics :37-60, mass :62-83, propulsion :245-300,
                                                flight_control :309-984, aerodynamics :986-1917
    <ref>/aircraft/f16/Engines/F100-PW-229.xml  turbine constants :3-16, thrust tables :26-82

Outputs (committed, because /root/reference does not exist on the GPU box):
    oracle/f16_oracle_gen.inc          tables + FCS channel wiring + aero function list as C++ for
                                       the CPU oracle (test infrastructure)
    f16_jsb_b200/csrc/f16_model_data.h named constants + tables for the CUDA kernels (product)
    tests/golden/f16_model.json        the parsed model as JSON (table known-answer tests)

The two C outputs are deliberately different renderings: the oracle gets a literal, component by
component transcription of the XML (every FCS component and every aero <function> in file order,
dead-ends included); the product header only gets numbers - the kernel's fused control laws and
coefficient build-up are hand-written in f16_model.cuh and are checked against the oracle.

Usage: python tools/gen_model.py [--ref /root/reference]
"""
import argparse
import json
import os
import re
import sys
import xml.etree.ElementTree as ET

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def mangle(prop: str) -> str:
    prop = prop.strip()
    return re.sub(r"[^A-Za-z0-9]", "_", prop)


def fnum(x) -> str:
    """Render a python float as a C double literal that round-trips."""
    r = repr(float(x))
    if "e" not in r and "." not in r and "inf" not in r and "nan" not in r:
        r += ".0"
    return r


def parse_table(tab):
    """Return dict(kind='1d'|'2d', rows, cols, data, indep=[(lookup, prop)])."""
    ivs = [(iv.get("lookup") or "row", iv.text.strip()) for iv in tab.findall("independentVar")]
    toks = tab.find("tableData").text.split()
    vals = [float(t) for t in toks]
    if len(ivs) == 1:
        rows = vals[0::2]
        data = vals[1::2]
        assert len(rows) == len(data)
        return dict(kind="1d", rows=rows, data=data, row_prop=ivs[0][1])
    assert len(ivs) == 2
    row_prop = [p for k, p in ivs if k == "row"][0]
    col_prop = [p for k, p in ivs if k == "column"][0]
    # first text line holds the column keys
    lines = [ln.split() for ln in tab.find("tableData").text.strip().splitlines() if ln.strip()]
    cols = [float(t) for t in lines[0]]
    rows, data = [], []
    for ln in lines[1:]:
        assert len(ln) == len(cols) + 1, ln
        rows.append(float(ln[0]))
        data.append([float(t) for t in ln[1:]])
    return dict(kind="2d", rows=rows, cols=cols, data=data, row_prop=row_prop, col_prop=col_prop)


# ----------------------------------------------------------------------------- model parse
def parse_model(ref):
    f16 = ET.parse(os.path.join(ref, "aircraft/f16/f16.xml")).getroot()
    eng = ET.parse(os.path.join(ref, "aircraft/f16/Engines/F100-PW-229.xml")).getroot()
    M = {}

    def triple(el):
        return [float(el.find(k).text) for k in ("x", "y", "z")]

    met = f16.find("metrics")
    M["metrics"] = dict(
        Sw=float(met.find("wingarea").text), bw=float(met.find("wingspan").text),
        cbar=float(met.find("chord").text),
        **{loc.get("name"): triple(loc) for loc in met.findall("location")})
    mb = f16.find("mass_balance")
    M["mass"] = dict(
        negated_crossproduct_inertia=mb.get("negated_crossproduct_inertia"),
        **{k: float(mb.find(k).text) for k in ("ixx", "iyy", "izz", "ixy", "ixz", "iyz", "emptywt")},
        cg=triple(mb.find("location")),
        pointmass=[dict(name=pm.get("name"), weight=float(pm.find("weight").text),
                        loc=triple(pm.find("location"))) for pm in mb.findall("pointmass")])
    prop = f16.find("propulsion")
    thr = prop.find("engine").find("thruster")
    M["propulsion"] = dict(
        thruster_loc=triple(thr.find("location")),
        thruster_orient=[float(thr.find("orient").find(k).text) for k in ("roll", "pitch", "yaw")],
        tanks=[dict(loc=triple(t.find("location")), capacity=float(t.find("capacity").text),
                    contents=float(t.find("contents").text)) for t in prop.findall("tank")])
    # ---- ground reactions (f16.xml:85-215): contact points, units as in the file (IN, LBS/FT, LBS/FT/SEC)
    contacts = []
    for c in f16.find("ground_reactions").findall("contact"):
        assert c.find("location").get("unit") == "IN"
        assert c.find("spring_coeff").get("unit") == "LBS/FT" and c.find("damping_coeff").get("unit") == "LBS/FT/SEC"
        assert c.find("damping_coeff_rebound") is None and c.find("orientation") is None
        contacts.append(dict(
            name=c.get("name"), type=c.get("type"), loc=triple(c.find("location")),
            static_friction=float(c.find("static_friction").text), dynamic_friction=float(c.find("dynamic_friction").text),
            rolling_friction=float(c.find("rolling_friction").text), spring=float(c.find("spring_coeff").text),
            damping=float(c.find("damping_coeff").text),
            retractable=int(float(c.find("retractable").text)) if c.find("retractable") is not None else 0))
    M["contacts"] = contacts
    E = {}
    for k in ("milthrust", "maxthrust", "bypassratio", "tsfc", "atsfc", "bleed", "idlen1", "idlen2",
              "maxn1", "maxn2", "augmented", "augmethod", "injected"):
        E[k] = float(eng.find(k).text)
    E["tables"] = {fn.get("name"): parse_table(fn.find("table")) for fn in eng.findall("function")}
    M["engine"] = E

    # ---- flight control channels, in file order
    chans = []
    for ch in f16.find("flight_control").findall("channel"):
        comps = []
        for c in ch:
            comps.append(parse_component(c))
        chans.append(dict(name=ch.get("name"), components=comps))
    M["fcs"] = chans

    # ---- aerodynamics
    aero = f16.find("aerodynamics")
    pre = []
    for fn in aero.findall("function"):
        assert len(list(fn)) == 2 and fn.find("table") is not None
        pre.append(dict(name=fn.get("name"), table=parse_table(fn.find("table"))))
    axes = []
    for ax in aero.findall("axis"):
        fns = []
        for fn in ax.findall("function"):
            prod = fn.find("product")
            assert prod is not None
            factors = []
            for c in prod:
                if c.tag == "property":
                    factors.append(dict(kind="property", prop=c.text.strip()))
                elif c.tag == "value":
                    factors.append(dict(kind="value", value=float(c.text)))
                elif c.tag == "table":
                    factors.append(dict(kind="table", table=parse_table(c)))
                else:
                    raise ValueError(c.tag)
            fns.append(dict(name=fn.get("name"), factors=factors))
        axes.append(dict(name=ax.get("name"), functions=fns))
    M["aero"] = dict(pre=pre, axes=axes)
    return M


def parse_clip(c):
    cl = c.find("clipto")
    if cl is None:
        return None
    return [float(cl.find("min").text), float(cl.find("max").text)]


def num_or_prop(s):
    s = s.strip()
    try:
        return float(s)
    except ValueError:
        return s


def parse_component(c):
    d = dict(type=c.tag, name=c.get("name"), clip=parse_clip(c),
             inputs=[i.text.strip() for i in c.findall("input")],
             outputs=[o.text.strip() for o in c.findall("output")])
    if c.tag == "switch":
        d["default"] = num_or_prop(c.find("default").get("value"))
        tests = []
        for t in c.findall("test"):
            conds = []
            for ln in t.text.strip().splitlines():
                ln = ln.strip()
                if not ln:
                    continue
                p, op, v = ln.split()
                conds.append((p, op.lower(), num_or_prop(v)))
            tests.append(dict(logic=(t.get("logic") or "AND").upper(), value=num_or_prop(t.get("value")),
                              conds=conds))
        d["tests"] = tests
    elif c.tag == "pure_gain":
        d["gain"] = float(c.find("gain").text)
    elif c.tag == "scheduled_gain":
        d["table"] = parse_table(c.find("table"))
    elif c.tag == "aerosurface_scale":
        dom = c.find("domain")
        rng = c.find("range")
        d["domain"] = [float(dom.find("min").text), float(dom.find("max").text)] if dom is not None else [-1.0, 1.0]
        d["range"] = [float(rng.find("min").text), float(rng.find("max").text)]
        assert c.find("zero_centered") is None and c.find("gain") is None
    elif c.tag == "summer":
        b = c.find("bias")
        d["bias"] = float(b.text) if b is not None else 0.0
    elif c.tag == "kinematic":
        d["detents"] = [(float(s.find("position").text), float(s.find("time").text))
                        for s in c.find("traverse").findall("setting")]
        assert c.find("noscale") is None
    elif c.tag == "pid":
        d["trigger"] = c.find("trigger").text.strip() if c.find("trigger") is not None else None
        for k in ("kp", "ki", "kd"):
            d[k] = float(c.find(k).text)
            assert c.find(k).get("type") is None
        assert c.get("type") is None  # not "standard"
    elif c.tag == "fcs_function":
        d["function"] = parse_expr(list(c.find("function"))[0])
    else:
        raise ValueError("unhandled FCS component " + c.tag)
    return d


def parse_expr(e):
    if e.tag == "property":
        return ("property", e.text.strip())
    if e.tag == "value":
        return ("value", float(e.text))
    if e.tag in ("product", "sum", "cos", "sin"):
        return (e.tag, [parse_expr(ch) for ch in e])
    raise ValueError(e.tag)


# ----------------------------------------------------------------------------- oracle codegen
class OracleGen:
    def __init__(self, M):
        self.M = M
        self.props = []          # ordered unique property names
        self.tables = []         # (cname, table)
        self.lines = []

    def P(self, prop):
        prop = prop.strip()
        if prop not in self.props:
            self.props.append(prop)
        return "P." + mangle(prop)

    def signed(self, inp):
        inp = inp.strip()
        if inp.startswith("-"):
            return "(-" + self.P(inp[1:]) + ")"
        return self.P(inp)

    def val(self, v):
        return fnum(v) if isinstance(v, float) else self.P(v)

    def add_table(self, hint, t):
        cname = "T_" + mangle(hint)
        self.tables.append((cname, t))
        return cname

    def table_call(self, cname, t):
        if t["kind"] == "1d":
            return "table1d(%s_x, %s_y, %d, %s)" % (cname, cname, len(t["rows"]), self.P(t["row_prop"]))
        return "table2d(%s_r, %s_c, &%s_v[0][0], %d, %d, %s, %s)" % (
            cname, cname, cname, len(t["rows"]), len(t["cols"]), self.P(t["row_prop"]), self.P(t["col_prop"]))

    def expr(self, e):
        k = e[0]
        if k == "property":
            return self.P(e[1])
        if k == "value":
            return fnum(e[1])
        if k == "product":
            return "(" + " * ".join(self.expr(x) for x in e[1]) + ")"
        if k == "sum":
            return "(" + " + ".join(self.expr(x) for x in e[1]) + ")"
        if k in ("cos", "sin"):
            return "std::%s(%s)" % (k, self.expr(e[1][0]))
        raise ValueError(k)

    def finish(self, c, out):
        """clip -> own property -> <output> properties (FGFCSComponent::Clip / SetOutput)."""
        L = out
        if c["clip"]:
            L.append("    o = constrain(%s, o, %s);" % (fnum(c["clip"][0]), fnum(c["clip"][1])))
        L.append("    %s = o;" % self.P(c["name"]))
        for o in c["outputs"]:
            L.append("    %s = o;" % self.P(o))

    def component(self, c, L, mem):
        t = c["type"]
        L.append("  { // <%s name=\"%s\">" % (t, c["name"]))
        L.append("    double o;")
        if t == "switch":
            L.append("    bool pass = false; o = 0.0;")
            ops = {"lt": "<", "le": "<=", "gt": ">", "ge": ">=", "eq": "==", "==": "==", "ne": "!="}
            for ts in c["tests"]:
                conds = ["(%s %s %s)" % (self.P(p), ops[op], self.val(v)) for p, op, v in ts["conds"]]
                j = " && " if ts["logic"] == "AND" else " || "
                L.append("    if (!pass && (%s)) { o = %s; pass = true; }" % (j.join(conds), self.val(ts["value"])))
            L.append("    if (!pass) o = %s;" % self.val(c["default"]))
        elif t == "pure_gain":
            L.append("    o = %s * %s;" % (fnum(c["gain"]), self.signed(c["inputs"][0])))
        elif t == "scheduled_gain":
            cn = self.add_table(c["name"], c["table"])
            L.append("    o = %s * %s;" % (self.table_call(cn, c["table"]), self.signed(c["inputs"][0])))
        elif t == "aerosurface_scale":
            L.append("    o = aerosurface_scale(%s, %s, %s, %s, %s);" % (
                self.signed(c["inputs"][0]), fnum(c["domain"][0]), fnum(c["domain"][1]),
                fnum(c["range"][0]), fnum(c["range"][1])))
        elif t == "summer":
            L.append("    o = 0.0;")
            for i in c["inputs"]:
                L.append("    o += %s;" % self.signed(i))
            L.append("    o += %s;" % fnum(c["bias"]))
        elif t == "kinematic":
            n = len(c["detents"])
            L.append("    static const double det[%d] = {%s};" % (n, ", ".join(fnum(d[0]) for d in c["detents"])))
            L.append("    static const double tim[%d] = {%s};" % (n, ", ".join(fnum(d[1]) for d in c["detents"])))
            cur = self.P(c["outputs"][0]) if c["outputs"] else self.P(c["name"])
            L.append("    o = kinematic_run(det, tim, %d, %s, %s, dt);" % (n, self.signed(c["inputs"][0]), cur))
        elif t == "pid":
            m = "M." + mangle(c["name"])
            mem.append(mangle(c["name"]))
            trig = self.P(c["trigger"]) if c["trigger"] else "0.0"
            L.append("    o = pid_run(%s, %s, %s, %s, %s, %s, dt);" % (
                m, self.signed(c["inputs"][0]), trig, fnum(c["kp"]), fnum(c["ki"]), fnum(c["kd"])))
        elif t == "fcs_function":
            L.append("    o = %s;" % self.expr(c["function"]))
        self.finish(c, L)
        L.append("  }")

    def generate(self):
        M = self.M
        fcs_lines, mem = [], []
        for ch in M["fcs"]:
            fcs_lines.append("  // ---- channel \"%s\"" % ch["name"])
            for c in ch["components"]:
                self.component(c, fcs_lines, mem)
        aero_lines = []
        for pf in M["aero"]["pre"]:
            cn = self.add_table(pf["name"], pf["table"])
            aero_lines.append("  %s = %s; // pre-function (FGModelFunctions::RunPreFunctions)" % (
                self.P(pf["name"]), self.table_call(cn, pf["table"])))
        axis_idx = {"DRAG": 0, "SIDE": 1, "LIFT": 2, "ROLL": 3, "PITCH": 4, "YAW": 5}
        nfun = 0
        for ax in M["aero"]["axes"]:
            ai = axis_idx[ax["name"]]
            aero_lines.append("  // ---- axis %s" % ax["name"])
            for fn in ax["functions"]:
                aero_lines.append("  { // %s" % fn["name"])
                first = True
                for f in fn["factors"]:
                    if f["kind"] == "property":
                        e = self.P(f["prop"])
                    elif f["kind"] == "value":
                        e = fnum(f["value"])
                    else:
                        cn = self.add_table(fn["name"], f["table"])
                        e = self.table_call(cn, f["table"])
                    aero_lines.append("    %s %s;" % ("double v =" if first else "v *=", e))
                    first = False
                aero_lines.append("    axis[%d] += v; fval[%d] = v;" % (ai, nfun))
                aero_lines.append("  }")
                nfun += 1

        out = []
        out.append("// GENERATED by tools/gen_model.py from the reference's aircraft/f16/f16.xml and")
        out.append("// aircraft/f16/Engines/F100-PW-229.xml - do not edit. TEST INFRASTRUCTURE (oracle) only.")
        out.append("// FCS components appear in file order (f16.xml:317-982), aero functions in file order")
        out.append("// (f16.xml:988-1915); every number below is copied from those files.")
        out.append("")
        out.append("struct Props {")
        for p in self.props:
            out.append("  double %s = 0.0; // %s" % (mangle(p), p))
        out.append("};")
        out.append("struct FcsMem {")
        for m in mem:
            out.append("  PidMem %s;" % m)
        out.append("};")
        out.append("static const char* const kPropNames[] = {%s};" % ", ".join('"%s"' % p for p in self.props))
        out.append("static const int kNumProps = %d;" % len(self.props))
        out.append("static const int kNumAeroFunctions = %d;" % nfun)
        fnames = [fn["name"] for ax in M["aero"]["axes"] for fn in ax["functions"]]
        out.append("static const char* const kAeroFunctionNames[] = {%s};" % ", ".join('"%s"' % n for n in fnames))
        out.append("")
        eng = M["engine"]
        for k, t in eng["tables"].items():
            self.tables.append(("T_ENG_" + k, t))
        for cname, t in self.tables:
            if t["kind"] == "1d":
                out.append("static const double %s_x[%d] = {%s};" % (cname, len(t["rows"]), ", ".join(map(fnum, t["rows"]))))
                out.append("static const double %s_y[%d] = {%s};" % (cname, len(t["data"]), ", ".join(map(fnum, t["data"]))))
            else:
                out.append("static const double %s_r[%d] = {%s};" % (cname, len(t["rows"]), ", ".join(map(fnum, t["rows"]))))
                out.append("static const double %s_c[%d] = {%s};" % (cname, len(t["cols"]), ", ".join(map(fnum, t["cols"]))))
                out.append("static const double %s_v[%d][%d] = {" % (cname, len(t["rows"]), len(t["cols"])))
                for r in t["data"]:
                    out.append("  {%s}," % ", ".join(map(fnum, r)))
                out.append("};")
        out.append("")
        out.append("// FGFCS::Run channel loop (f16.xml:317-982), one block per component in file order")
        out.append("static void fcs_channels_run(Props& P, FcsMem& M, double dt) {")
        out.extend(fcs_lines)
        out.append("}")
        out.append("")
        out.append("// FGAerodynamics::Run function loop (f16.xml:988-1915): axis[0..2] = DRAG, SIDE, LIFT; axis[3..5] = ROLL, PITCH, YAW")
        out.append("static void aero_functions_run(Props& P, double axis[6], double* fval) {")
        out.append("  for (int i = 0; i < 6; ++i) axis[i] = 0.0;")
        out.extend(aero_lines)
        out.append("}")
        out.append("")
        # scalar model constants
        m = M["metrics"]
        mb = M["mass"]
        pr = M["propulsion"]
        out.append("namespace modelk {")
        out.append("static const double Sw = %s, bw = %s, cbar = %s;" % (fnum(m["Sw"]), fnum(m["bw"]), fnum(m["cbar"])))
        for nm in ("AERORP", "EYEPOINT", "VRP"):
            out.append("static const double %s[3] = {%s};" % (nm, ", ".join(map(fnum, m[nm]))))
        out.append("static const bool negated_crossproduct_inertia = %s;" % ("true" if mb["negated_crossproduct_inertia"] != "false" else "false"))
        for k in ("ixx", "iyy", "izz", "ixy", "ixz", "iyz", "emptywt"):
            out.append("static const double %s = %s;" % (k, fnum(mb[k])))
        out.append("static const double base_cg[3] = {%s};" % ", ".join(map(fnum, mb["cg"])))
        out.append("static const int n_pointmass = %d;" % len(mb["pointmass"]))
        out.append("static const double pointmass_w[%d] = {%s};" % (len(mb["pointmass"]), ", ".join(fnum(p["weight"]) for p in mb["pointmass"])))
        out.append("static const double pointmass_loc[%d][3] = {%s};" % (len(mb["pointmass"]), ", ".join("{%s}" % ", ".join(map(fnum, p["loc"])) for p in mb["pointmass"])))
        nt = len(pr["tanks"])
        out.append("static const int n_tanks = %d;" % nt)
        out.append("static const double tank_loc[%d][3] = {%s};" % (nt, ", ".join("{%s}" % ", ".join(map(fnum, t["loc"])) for t in pr["tanks"])))
        out.append("static const double tank_capacity[%d] = {%s};" % (nt, ", ".join(fnum(t["capacity"]) for t in pr["tanks"])))
        out.append("static const double tank_contents0[%d] = {%s};" % (nt, ", ".join(fnum(t["contents"]) for t in pr["tanks"])))
        out.append("static const double thruster_loc[3] = {%s};" % ", ".join(map(fnum, pr["thruster_loc"])))
        for k in ("milthrust", "maxthrust", "bypassratio", "tsfc", "atsfc", "bleed", "idlen1", "idlen2", "maxn1", "maxn2"):
            out.append("static const double %s = %s;" % (k, fnum(eng[k])))
        for k in ("augmented", "augmethod", "injected"):
            out.append("static const int %s = %d;" % (k, int(eng[k])))
        cs = M["contacts"]
        out.append("// ground_reactions f16.xml:85-215, file order")
        out.append("static const int n_contacts = %d;" % len(cs))
        out.append("struct ContactDef { const char* name; bool bogey; bool retractable; double loc[3]; double static_f, dynamic_f, rolling_f, spring, damping; };")
        out.append("static const ContactDef contacts[%d] = {" % len(cs))
        for c in cs:
            out.append('  {"%s", %s, %s, {%s}, %s, %s, %s, %s, %s},' % (
                c["name"], "true" if c["type"] == "BOGEY" else "false", "true" if c["retractable"] else "false",
                ", ".join(map(fnum, c["loc"])), fnum(c["static_friction"]), fnum(c["dynamic_friction"]),
                fnum(c["rolling_friction"]), fnum(c["spring"]), fnum(c["damping"])))
        out.append("};")
        out.append("} // namespace modelk")
        return "\n".join(out) + "\n"


# ----------------------------------------------------------------------------- product header
def find_comp(M, name):
    for ch in M["fcs"]:
        for c in ch["components"]:
            if c["name"] == name:
                return c
    raise KeyError(name)


def aero_fn(M, name):
    for ax in M["aero"]["axes"]:
        for fn in ax["functions"]:
            if fn["name"].endswith("/" + name):
                return fn
    raise KeyError(name)


def fn_table(fn):
    return [f["table"] for f in fn["factors"] if f["kind"] == "table"][0]


def fn_value(fn):
    return [f["value"] for f in fn["factors"] if f["kind"] == "value"][0]


def gen_product_header(M):
    """Numbers only. The layout is chosen for the kernel (see f16_model.cuh):

    * all 16 one-D alpha tables interleaved alpha-major:  A1[12][16]
    * the three alpha x elevator tables interleaved:      AE[12][5][4]   (CD, CL, Cm, pad)
    * the two alpha x beta13 tables interleaved:          AB13[12][13][2] (Cl, Cn)
    * the four alpha x beta7 tables interleaved:          AB7[12][7][4]  (Clda, Cldr, Cnda, Cndr)
    so that one (index, fraction) pair per independent variable fetches every coefficient with
    vector loads from shared memory.
    """
    o = []
    o.append("// GENERATED by tools/gen_model.py from the reference's aircraft/f16/f16.xml and")
    o.append("// aircraft/f16/Engines/F100-PW-229.xml - do not edit. Numbers only; kernel layout.")
    o.append("#pragma once")
    o.append("namespace f16data {")
    m, mb, pr, eng = M["metrics"], M["mass"], M["propulsion"], M["engine"]
    o.append("// metrics f16.xml:37-60")
    o.append("constexpr double Sw = %s, bw = %s, cbar = %s;" % (fnum(m["Sw"]), fnum(m["bw"]), fnum(m["cbar"])))
    for nm in ("AERORP", "EYEPOINT"):
        o.append("constexpr double %s[3] = {%s};" % (nm, ", ".join(map(fnum, m[nm]))))
    o.append("// mass_balance f16.xml:62-83 (negated_crossproduct_inertia=\"%s\")" % mb["negated_crossproduct_inertia"])
    o.append("constexpr bool negated_crossproduct_inertia = %s;" % ("true" if mb["negated_crossproduct_inertia"] != "false" else "false"))
    for k in ("ixx", "iyy", "izz", "ixy", "ixz", "iyz", "emptywt"):
        o.append("constexpr double %s = %s;" % (k, fnum(mb[k])))
    o.append("constexpr double base_cg[3] = {%s};" % ", ".join(map(fnum, mb["cg"])))
    assert len(mb["pointmass"]) == 1
    o.append("constexpr double pilot_w = %s;" % fnum(mb["pointmass"][0]["weight"]))
    o.append("constexpr double pilot_loc[3] = {%s};" % ", ".join(map(fnum, mb["pointmass"][0]["loc"])))
    o.append("// propulsion f16.xml:245-300")
    nt = len(pr["tanks"])
    o.append("constexpr int n_tanks = %d;" % nt)
    o.append("constexpr double tank_loc[%d][3] = {%s};" % (nt, ", ".join("{%s}" % ", ".join(map(fnum, t["loc"])) for t in pr["tanks"])))
    o.append("constexpr double tank_contents0[%d] = {%s};" % (nt, ", ".join(fnum(t["contents"]) for t in pr["tanks"])))
    o.append("constexpr double thruster_loc[3] = {%s};" % ", ".join(map(fnum, pr["thruster_loc"])))
    o.append("// Engines/F100-PW-229.xml:3-16")
    for k in ("milthrust", "maxthrust", "bypassratio", "bleed", "idlen1", "idlen2", "maxn1", "maxn2"):
        o.append("constexpr double %s = %s;" % (k, fnum(eng[k])))
    assert int(eng["augmethod"]) == 2 and int(eng["augmented"]) == 1 and int(eng["injected"]) == 0

    # ---- structure contacts (f16.xml:137-214); the three BOGEY contacts are retractable and the env keeps the gear up
    st = [c for c in M["contacts"] if c["type"] == "STRUCTURE"]
    assert all(c["retractable"] for c in M["contacts"] if c["type"] == "BOGEY")
    o.append("// ground_reactions f16.xml:137-214: STRUCTURE contacts (location IN, spring LBS/FT, damping LBS/FT/SEC)")
    o.append("constexpr int n_structure = %d;" % len(st))
    o.append("#define F16_STRUCT_LOC {%s}" % ", ".join("{%s}" % ", ".join(map(fnum, c["loc"])) for c in st))
    o.append("#define F16_STRUCT_SPRING {%s}" % ", ".join(fnum(c["spring"]) for c in st))
    o.append("#define F16_STRUCT_DAMPING {%s}" % ", ".join(fnum(c["damping"]) for c in st))
    o.append("#define F16_STRUCT_STATIC_F {%s}" % ", ".join(fnum(c["static_friction"]) for c in st))
    o.append("#define F16_STRUCT_DYNAMIC_F {%s}" % ", ".join(fnum(c["dynamic_friction"]) for c in st))
    # ---- FCS named constants (f16.xml:317-935)
    o.append("// flight_control f16.xml:317-935")
    sw = find_comp(M, "fcs/tef-pos-rad")
    o.append("constexpr double tef_lowspeed_rad = %s, tef_vc_kts = %s, tef_highmach_rad = %s, tef_mach = %s;" % (
        fnum(sw["tests"][0]["value"]), fnum(sw["tests"][0]["conds"][0][2]),
        fnum(sw["tests"][1]["value"]), fnum(sw["tests"][1]["conds"][0][2])))
    o.append("constexpr double tef_norm_gain = %s;" % fnum(find_comp(M, "fcs/tef-pos-norm")["gain"]))
    k = find_comp(M, "fcs/tef-control")
    assert [d[0] for d in k["detents"]] == [-1.0, 0.0, 1.0] and k["detents"][1][1] == 0.0
    o.append("constexpr double tef_time_pos = %s;" % fnum(k["detents"][2][1]))
    o.append("constexpr double roll_rate_gain = %s;" % fnum(find_comp(M, "fcs/roll-rate-norm")["gain"]))
    for nm, comp in (("roll", "fcs/roll-rate-pid"), ("pitch", "fcs/g-load-pid"), ("yaw", "fcs/yaw-load-pid")):
        c = find_comp(M, comp)
        o.append("constexpr double %s_kp = %s, %s_ki = %s, %s_kd = %s;" % (nm, fnum(c["kp"]), nm, fnum(c["ki"]), nm, fnum(c["kd"])))
    o.append("constexpr double roll_trigger_kts = %s, pitch_trigger_kts = %s, yaw_trigger_kts = %s;" % (
        fnum(find_comp(M, "fcs/aileron-pid-trigger")["tests"][0]["conds"][0][2]),
        fnum(find_comp(M, "fcs/elevator-pid-trigger")["tests"][0]["conds"][0][2]),
        fnum(find_comp(M, "fcs/rudder-pid-trigger")["tests"][0]["conds"][0][2])))
    o.append("constexpr double aileron_max_rad = %s;" % fnum(find_comp(M, "fcs/aileron-control")["range"][1]))
    assert find_comp(M, "fcs/aileron-control")["range"][0] == -find_comp(M, "fcs/aileron-control")["range"][1]
    for nm, comp in (("aileron", "fcs/aileron-position"), ("elevator", "fcs/elevator-position-normalized"),
                     ("rudder", "fcs/rudder-position")):
        k = find_comp(M, comp)
        assert [d[0] for d in k["detents"]] == [-1.0, 1.0]
        o.append("constexpr double %s_traverse_s = %s;" % (nm, fnum(k["detents"][1][1])))
    def small_table(nm, t):
        # macro initialiser lists: device code builds a local array from them (namespace-scope
        # constexpr arrays cannot be indexed from device code)
        o.append("constexpr int %s_n = %d;" % (nm, len(t["rows"])))
        o.append("#define F16_%s_X {%s}" % (nm.upper(), ", ".join(map(fnum, t["rows"]))))
        o.append("#define F16_%s_Y {%s}" % (nm.upper(), ", ".join(map(fnum, t["data"]))))
    small_table("ail_comp", find_comp(M, "fcs/aileron-speed-compensated")["table"])
    o.append("constexpr double flaperon_mix_gain = %s;" % fnum(find_comp(M, "fcs/flaperon-mix-rad")["gain"]))
    c = find_comp(M, "fcs/elevator-cmd-limiter")
    o.append("constexpr double elev_cmd_min = %s, elev_cmd_max = %s;" % (fnum(c["clip"][0]), fnum(c["clip"][1])))
    small_table("elev_sched", find_comp(M, "fcs/elevator-scheduler")["table"])
    o.append("constexpr double alpha_limiter_gain = %s, pitch_rate_gain = %s, g_load_gain = %s;" % (
        fnum(find_comp(M, "fcs/alpha-limiter-norm")["gain"]), fnum(find_comp(M, "fcs/pitch-rate-norm")["gain"]),
        fnum(find_comp(M, "fcs/g-load-norm")["gain"])))
    o.append("constexpr double elevator_max_rad = %s;" % fnum(find_comp(M, "fcs/elevator-position")["range"][1]))
    small_table("yaw_rate", find_comp(M, "fcs/yaw-rate-norm")["table"])
    o.append("constexpr double yaw_load_gain = %s;" % fnum(find_comp(M, "fcs/yaw-load-norm")["gain"]))
    o.append("constexpr double rudder_max_rad = %s;" % fnum(find_comp(M, "fcs/rudder-control")["range"][1]))
    k = find_comp(M, "fcs/gear-control")
    o.append("constexpr double gear_traverse_s = %s;" % fnum(k["detents"][1][1]))
    sw = find_comp(M, "fcs/lef-pos-rad")
    o.append("constexpr double lef_ground_rad = %s, lef_hi_alpha = %s, lef_hi_rad = %s, lef_mid_alpha = %s, lef_mid_rad = %s, lef_mach = %s, lef_mach_rad = %s;" % (
        fnum(sw["tests"][0]["value"]), fnum(sw["tests"][1]["conds"][1][2]), fnum(sw["tests"][1]["value"]),
        fnum(sw["tests"][2]["conds"][1][2]), fnum(sw["tests"][2]["value"]),
        fnum(sw["tests"][3]["conds"][0][2]), fnum(sw["tests"][3]["value"])))
    o.append("constexpr double throttle_gain = %s;" % fnum(find_comp(M, "fcs/throttle1")["gain"]))
    sw = find_comp(M, "fcs/speedbrake-alpha-limiter")
    o.append("constexpr double sb_alpha_deg = %s, sb_v_fps = %s;" % (fnum(sw["tests"][0]["conds"][0][2]), fnum(sw["tests"][0]["conds"][1][2])))
    small_table("sb_sched", find_comp(M, "fcs/speedbrake-scheduler")["table"])
    k = find_comp(M, "fcs/speedbrake-control")
    assert k["detents"][0] == (0.0, 0.0)
    o.append("constexpr double sb_max_deg = %s, sb_traverse_s = %s;" % (fnum(k["detents"][1][0]), fnum(k["detents"][1][1])))

    # ---- aero scalar factors
    o.append("// aerodynamics f16.xml:986-1917 (scalar factors)")
    for nm in ("CDDflaps", "CDgear", "CYb", "CYDa", "CYdr", "CLDflaps"):
        o.append("constexpr double k_%s = %s;" % (nm, fnum(fn_value(aero_fn(M, nm)))))

    # ---- tables
    alpha = fn_table(aero_fn(M, "CDDlef"))["rows"]
    NA = len(alpha)
    one_d = ["CDDlef", "CDDsb", "CDq", "CDq_Dlef", "CYp", "CYr", "CLDlef", "CLDsb", "CLq", "CLq_Dsb",
             "Clp", "Clr", "CmDsb", "Cmq", "Cnp", "Cnr"]
    for nm in one_d:
        t = fn_table(aero_fn(M, nm))
        assert t["kind"] == "1d" and t["rows"] == alpha and t["row_prop"] == "aero/alpha-rad", nm
    o.append("#define F16_ALPHA_BP {%s}" % ", ".join(map(fnum, alpha)))
    o.append("constexpr int NA = %d;  // alpha breakpoints shared by all alpha-indexed tables" % NA)
    o.append("constexpr double alpha_bp[NA] = {%s};" % ", ".join(map(fnum, alpha)))
    o.append("// column order of A1: " + ", ".join("%d=%s" % (i, n) for i, n in enumerate(one_d)))
    o.append("enum { " + ", ".join("A1_%s = %d" % (n, i) for i, n in enumerate(one_d)) + ", A1_N = %d };" % len(one_d))
    o.append("constexpr double A1[NA][A1_N] = {")
    for i in range(NA):
        o.append("  {%s}," % ", ".join(fnum(fn_table(aero_fn(M, nm))["data"][i]) for nm in one_d))
    o.append("};")
    # alpha x elevator
    ae = ["CDDh", "CLDh", "CmDh"]
    de = fn_table(aero_fn(M, "CDDh"))["cols"]
    for nm in ae:
        t = fn_table(aero_fn(M, nm))
        assert t["rows"] == alpha and t["cols"] == de and t["col_prop"] == "fcs/elevator-pos-rad", nm
    o.append("#define F16_DE_BP {%s}" % ", ".join(map(fnum, de)))
    o.append("constexpr int NDE = %d;" % len(de))
    o.append("constexpr double de_bp[NDE] = {%s};" % ", ".join(map(fnum, de)))
    o.append("// AE[alpha][de][k], k: 0=CDDh 1=CLDh 2=CmDh 3=pad")
    o.append("constexpr double AE[NA][NDE][4] = {")
    for i in range(NA):
        o.append("  {%s}," % ", ".join("{%s, 0.0}" % ", ".join(fnum(fn_table(aero_fn(M, nm))["data"][i][j]) for nm in ae) for j in range(len(de))))
    o.append("};")
    # alpha x beta13
    b13n = ["Clb", "Cnb"]
    b13 = fn_table(aero_fn(M, "Clb"))["cols"]
    for nm in b13n:
        t = fn_table(aero_fn(M, nm))
        assert t["rows"] == alpha and t["cols"] == b13 and t["col_prop"] == "aero/beta-rad", nm
    o.append("#define F16_B13_BP {%s}" % ", ".join(map(fnum, b13)))
    o.append("constexpr int NB13 = %d;" % len(b13))
    o.append("constexpr double b13_bp[NB13] = {%s};" % ", ".join(map(fnum, b13)))
    o.append("// AB13[alpha][beta][k], k: 0=Clb 1=Cnb")
    o.append("constexpr double AB13[NA][NB13][2] = {")
    for i in range(NA):
        o.append("  {%s}," % ", ".join("{%s}" % ", ".join(fnum(fn_table(aero_fn(M, nm))["data"][i][j]) for nm in b13n) for j in range(len(b13))))
    o.append("};")
    b7n = ["Clda", "Cldr", "Cnda", "Cndr"]
    b7 = fn_table(aero_fn(M, "Clda"))["cols"]
    for nm in b7n:
        t = fn_table(aero_fn(M, nm))
        assert t["rows"] == alpha and t["cols"] == b7 and t["col_prop"] == "aero/beta-rad", nm
    o.append("#define F16_B7_BP {%s}" % ", ".join(map(fnum, b7)))
    o.append("constexpr int NB7 = %d;" % len(b7))
    o.append("constexpr double b7_bp[NB7] = {%s};" % ", ".join(map(fnum, b7)))
    o.append("// AB7[alpha][beta][k], k: 0=Clda 1=Cldr 2=Cnda 3=Cndr")
    o.append("constexpr double AB7[NA][NB7][4] = {")
    for i in range(NA):
        o.append("  {%s}," % ", ".join("{%s}" % ", ".join(fnum(fn_table(aero_fn(M, nm))["data"][i][j]) for nm in b7n) for j in range(len(b7))))
    o.append("};")
    # mach tables: piecewise linear, few breakpoints each
    mach = ["CDmach", "CYb_M", "Clb_M", "Clda_M", "Cldr_M", "Cma_M", "Cnb_M", "Cnda_M", "Cndr_M"]
    union = sorted(set(x for nm in mach for x in fn_table(aero_fn(M, nm))["rows"]))
    o.append("// union of the Mach breakpoints of the nine Mach tables (the kernel resamples them on this grid)")
    o.append("constexpr int NMACH_UNION = %d;" % len(union))
    o.append("#define F16_MACH_BP {%s}" % ", ".join(map(fnum, union)))
    o.append("// Mach-indexed 1-D tables")
    for nm in mach:
        t = fn_table(aero_fn(M, nm))
        assert t["kind"] == "1d" and t["row_prop"] == "velocities/mach"
        o.append("constexpr int n_%s = %d;" % (nm, len(t["rows"])))
        o.append("constexpr double x_%s[%d] = {%s}, y_%s[%d] = {%s};" % (
            nm, len(t["rows"]), ", ".join(map(fnum, t["rows"])), nm, len(t["rows"]), ", ".join(map(fnum, t["data"]))))
    t = M["aero"]["pre"][0]["table"]
    assert M["aero"]["pre"][0]["name"].endswith("kCLge")
    o.append("constexpr int n_kCLge = %d;" % len(t["rows"]))
    o.append("constexpr double x_kCLge[%d] = {%s}, y_kCLge[%d] = {%s};" % (
        len(t["rows"]), ", ".join(map(fnum, t["rows"])), len(t["rows"]), ", ".join(map(fnum, t["data"]))))
    # engine tables: Mach rows x density-altitude columns; columns shared
    et = eng["tables"]
    cols = et["IdleThrust"]["cols"]
    assert et["MilThrust"]["cols"] == cols and et["AugThrust"]["cols"] == cols
    o.append("// Engines/F100-PW-229.xml:26-82 thrust tables [mach][density altitude ft]")
    o.append("constexpr int NEH = %d;" % len(cols))
    o.append("constexpr double eng_alt_bp[NEH] = {%s};" % ", ".join(map(fnum, cols)))
    for k, nm in (("IdleThrust", "idle"), ("MilThrust", "mil"), ("AugThrust", "aug")):
        t = et[k]
        assert t["row_prop"] == "velocities/mach" and t["col_prop"] == "atmosphere/density-altitude"
        o.append("constexpr int n_%s_mach = %d;" % (nm, len(t["rows"])))
        o.append("constexpr double %s_mach_bp[%d] = {%s};" % (nm, len(t["rows"]), ", ".join(map(fnum, t["rows"]))))
        o.append("constexpr double %s_tbl[%d][NEH] = {" % (nm, len(t["rows"])))
        for r in t["data"]:
            o.append("  {%s}," % ", ".join(map(fnum, r)))
        o.append("};")
    o.append("} // namespace f16data")
    return "\n".join(o) + "\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    a = ap.parse_args()
    M = parse_model(a.ref)
    g = OracleGen(M)
    with open(os.path.join(ROOT, "oracle/f16_oracle_gen.inc"), "w") as f:
        f.write(g.generate())
    with open(os.path.join(ROOT, "f16_jsb_b200/csrc/f16_model_data.h"), "w") as f:
        f.write(gen_product_header(M))
    with open(os.path.join(ROOT, "tests/golden/f16_model.json"), "w") as f:
        json.dump(M, f, indent=1)
    print("props:", len(g.props), "tables:", len(g.tables))


if __name__ == "__main__":
    sys.exit(main())
