"""helpers shared by the GPU parity tests: hand an oracle-built operator to the CUDA engine"""
import numpy as np

from openems_b200 import Operator_CUDA


def operator_from_oracle(s, include=("exc", "mur", "upml", "lorentz", "sheets", "tfsf")):
    """copies the host-side operator data of an OracleSim into an Operator_CUDA, exactly the
    data Operator_CUDA::CreateEngine would read from the reference's Operator/Operator_Ext_*"""
    op = Operator_CUDA(s.N)
    op.SetOperatorArrays(s.coeff("vv"), s.coeff("vi"), s.coeff("ii"), s.coeff("iv"))
    op.SetTimestep(s.dT)
    op.SetMesh(s.x, s.y, s.z, s.grid_delta)
    sv, si, per = s.signal()
    op.SetExcitationSignal(sv, si, per)
    if "exc" in include:
        for w in (0, 1):
            idx, d, amp, delay = s.excitation(w)
            op.SetExcitation(w, idx, d, amp, delay)
    if "mur" in include:
        for m in s.mur_planes():
            op.AddMur(m["ny"], m["line"], m["shift"], m["coeff_nyP"], m["coeff_nyPP"], m["start_ts"])
    if "upml" in include:
        for b in s.upml_boxes():
            op.AddUPML(b["start"], b["n"], b["vv"], b["vvfn"], b["vvfo"], b["ii"], b["iifn"], b["iifo"])
    if "tfsf" in include and s.tfsf() is not None:
        t = s.tfsf()
        op.SetTFSF(t["start"], t["stop"], t["active"], t["faces"])
    if "sheets" in include:
        for a in s.absorbing_sheets():
            sa = a["type"] == 2
            op.AddAbsorbingSheet(a["ny"], a["x0"], a["x1"], a["positive"], a["type"], a["K1P"], a["K1PP"],
                                 a["K2P"] if sa else None, a["K2PP"] if sa else None)
    if "lorentz" in include:
        for L in s.lorentz():
            op.AddLorentzOrder(L["pos"], L["v_int"], L["v_ext"], L["v_lor"], L["i_int"], L["i_ext"], L["i_lor"])
    return op


def assert_fields_equal(eng, s, what=""):
    """bit-exact comparison of E and H (TESTSUITE/enginetests/cavity.m:155 rule)"""
    v = eng.GetFields(0)
    c = eng.GetFields(1)
    sv, sc = s.volt, s.curr
    nv = int((v.view(np.uint32) != sv.view(np.uint32)).sum())
    nc = int((c.view(np.uint32) != sc.view(np.uint32)).sum())
    if nv or nc:
        d = np.abs(v.astype(np.float64) - sv).max(), np.abs(c.astype(np.float64) - sc).max()
        bad = np.argwhere(v.view(np.uint32) != sv.view(np.uint32))[:5].tolist()
        badc = np.argwhere(c.view(np.uint32) != sc.view(np.uint32))[:5].tolist()
        raise AssertionError("%s fields differ: %d volt / %d curr values, max abs diff %g / %g, first volt %s curr %s"
                             % (what, nv, nc, d[0], d[1], bad, badc))
    return float(np.abs(sv).max()), float(np.abs(sc).max())
