"""Oracle: mirror Fresnel amplitude and thick-Bragg crystal amplitude (TEST INFRASTRUCTURE ONLY).

Restates
  * Element.get_f1f2 / get_f0            xrt/backends/raycing/materials/element.py:252-263, :203-207
  * Material.get_refractive_index        materials/material.py:348-378
  * Material.get_amplitude               materials/material.py:415-493 (kinds 'mirror', 'thin mirror')
  * Crystal.get_amplitude (thick Bragg)  materials/crystal.py:492-645 (branch :568-597)
  * Crystal.get_F_chi                    materials/crystal.py:297-306
  * CrystalFcc/CrystalDiamond.get_structure_factor   materials/crystals_basic.py:22-31, :76-80
  * Crystal.get_Bragg_angle              materials/crystal.py:1105-1120
The element tables come from xrt_b200/data/elements.npz (a verbatim subset of the reference's data files,
see tools/extract_tables.py)."""
import os

import numpy as np

from .physconsts import PI, PI2, CH, CHBAR, R0, AVOGADRO

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'xrt_b200', 'data',
                     'elements.npz')
_tables = None


def _load():
    global _tables
    if _tables is None:
        with np.load(_DATA) as f:
            _tables = {k: np.array(f[k]) for k in f.files}
    return _tables


class Element(object):
    def __init__(self, name, table='Chantler'):
        t = _load()
        self.name = name
        self.Z = int(t[name + '_Z'])
        self.mass = float(t[name + '_mass'])
        self.f0coeffs = [float(v) for v in t[name + '_f0']]
        f2key = '_f2tot' if 'total' in table else '_f2'  # element.py:212
        self.E, self.f1, self.f2 = t[name + '_E'], t[name + '_f1'], t[name + f2key]

    def get_f0(self, qOver4pi=0):
        return self.f0coeffs[5] + sum(a * np.exp(-b * qOver4pi**2)
                                      for a, b in zip(self.f0coeffs[:5], self.f0coeffs[6:]))

    def get_f1f2(self, E):
        if np.any(E < self.E[0]) or np.any(E > self.E[-1]):
            raise ValueError('E is out of the data table range')
        f1 = np.interp(E, self.E, self.f1)
        f2 = np.interp(E, self.E, self.f2)
        return f1 + 1j*f2


class EmptyMaterial(object):
    """materials/__init__.py:101-113: a 'grating' material without reflectivity (no get_amplitude)."""

    def __init__(self, kind='grating'):
        self.kind = kind
        self.geom = ''
        self.efficiency = None


class Material(object):
    def __init__(self, elements, quantities=None, kind='mirror', rho=0, t=None, table='Chantler total',
                 efficiency=None):
        self.efficiency = efficiency  # material.py:78-93, constant [(order, value), ...] only
        if isinstance(elements, str):
            elements = elements,
        self.elements = [Element(e, table) for e in elements]
        self.quantities = [1. for e in self.elements] if quantities is None else list(quantities)
        self.kind = kind
        self.rho = rho
        self.t = t
        self.mass = 0.
        for elem, xi in zip(self.elements, self.quantities):
            self.mass += xi * elem.mass

    def get_refractive_index(self, E):
        xf = np.zeros_like(E) * 0j
        for elem, xi in zip(self.elements, self.quantities):
            xf += (elem.Z + elem.get_f1f2(E)) * xi
        return 1 - 1e-24 * AVOGADRO * R0 / PI2 * (CH/E)**2 * self.rho * xf / self.mass

    def get_grating_efficiency(self, beam, good):
        """material.py:391-413 without an efficiency file."""
        resI = np.zeros(good.sum())
        order = beam.order[good]
        for eff in self.efficiency:
            resI[order == eff[0]] = eff[1]
        resA = resI**0.5
        return resA, resA, 0

    def get_amplitude(self, E, beamInDotNormal, fromVacuum=True):
        n = self.get_refractive_index(E)
        if fromVacuum:
            n1 = 1.
            n2 = n
        else:
            n1 = n
            n2 = 1.
        cosAlpha = abs(beamInDotNormal)
        sinAlpha2 = 1 - beamInDotNormal**2
        if isinstance(sinAlpha2, np.ndarray):
            sinAlpha2[sinAlpha2 < 0] = 0
        n1cosAlpha = n1 * cosAlpha
        cosBeta = np.sqrt(1 - (n1/n2)**2*sinAlpha2)
        n2cosBeta = n2 * cosBeta
        if self.kind in ('mirror', 'thin mirror', 'grating'):  # reflectivity
            rs = (n1cosAlpha - n2cosBeta) / (n1cosAlpha + n2cosBeta)
            rp = (n2*cosAlpha - n1*cosBeta) / (n2*cosAlpha + n1*cosBeta)
            if self.kind == 'thin mirror':
                p2 = np.exp(2j * E / CHBAR * n2cosBeta * self.t * 1e7)
                rs *= (1 - p2) / (1 - rs**2*p2)
                rp *= (1 - p2) / (1 - rp**2*p2)
        elif self.kind in ('plate', 'lens'):  # transmittivity (material.py:482-486)
            tf = np.sqrt((n2cosBeta * n1.conjugate()).real / cosAlpha) / abs(n1)
            rs = 2 * n1cosAlpha / (n1cosAlpha + n2cosBeta) * tf
            rp = 2 * n1cosAlpha / (n2*cosAlpha + n1*cosBeta) * tf
        else:
            raise ValueError('Unknown kind of material')
        return (rs, rp, abs(n.imag) * E / CHBAR * 2e8, n.real * E / CHBAR * 1e8)


class CrystalDiamond(object):
    """Single-element fcc/diamond crystal: thick Bragg reflection (t=None) or thickness t (mm) with geom
    Bragg / Laue, reflected / transmitted (crystal.py:566-616)."""
    kind = 'crystal'

    def __init__(self, hkl, d, element='Si', factDW=1., structure='diamond', table='Chantler total', t=None,
                 geom='Bragg reflected'):
        self.t = t
        self.geom = geom
        self.hkl = tuple(int(i) for i in hkl)
        self.d = d
        self.sqrthkl2 = (sum(i**2 for i in self.hkl))**0.5
        self.V = (d * self.sqrthkl2)**3
        self.chiToF = -R0 / PI / self.V
        self.factDW = factDW
        self.structure = structure
        self.elements = [Element(element, table)]

    def get_structure_factor(self, E, sinThetaOverLambda=0):
        anomalousPart = self.elements[0].get_f1f2(E)
        F0 = 4 * (self.elements[0].Z+anomalousPart) * self.factDW
        residue = sum(i % 2 for i in self.hkl)
        if residue == 0 or residue == 3:
            f0 = self.elements[0].get_f0(sinThetaOverLambda)
            Fhkl = 4 * (f0+anomalousPart) * self.factDW
        else:
            Fhkl = 0.
        if self.structure == 'fcc':
            return F0, Fhkl, Fhkl
        diamondToFcc = 1 + np.exp(0.5j * PI * sum(self.hkl))
        return F0 * 2, Fhkl * diamondToFcc, Fhkl * diamondToFcc.conjugate()

    def get_F_chi(self, E, sinThetaOverLambda):
        F0, Fhkl, Fhkl_ = self.get_structure_factor(E, sinThetaOverLambda)
        waveLength = CH / E
        lambdaSquare = waveLength**2
        chiToFlambdaSquare = self.chiToF * lambdaSquare
        chi0 = F0.conjugate() * chiToFlambdaSquare
        chih = Fhkl.conjugate() * chiToFlambdaSquare
        chih_ = Fhkl_.conjugate() * chiToFlambdaSquare
        return F0, Fhkl, Fhkl_, chi0, chih, chih_

    def get_Bragg_angle(self, E, order=1):
        a = order * CH / (2*self.d*E)
        a = np.array(a)
        a[a > 1] = 1 - 1e-16
        a[a < -1] = -1 + 1e-16
        return np.arcsin(a)

    def get_dtheta_symmetric_Bragg(self, E):
        F0, Fhkl, Fhkl_ = self.get_structure_factor(E, 0.5 / self.d)
        waveLength = CH / E
        lambdaSquare = waveLength ** 2
        chiToFlambdaSquare = self.chiToF * lambdaSquare
        chi0 = F0 * chiToFlambdaSquare
        thetaB = self.get_Bragg_angle(E)
        return (chi0 / np.sin(2*thetaB)).real

    def get_dtheta(self, E, alpha=None):
        """materials/crystal.py:1141-1173 (Bragg geometry)."""
        if alpha is None:
            alpha = 0
        thetaB = self.get_Bragg_angle(E)
        pm = -1
        gamma0 = np.sin(thetaB + alpha)
        gammah = pm * np.sin(thetaB - alpha)
        symm_dt = self.get_dtheta_symmetric_Bragg(E)
        osqg0 = np.sqrt(1. - gamma0**2)
        dtheta0 = (pm*gamma0 - pm*np.sqrt(gamma0**2 + pm*(gamma0 - gammah) * osqg0 * symm_dt)) / osqg0
        return -dtheta0

    def get_amplitude(self, E, beamInDotNormal, beamOutDotNormal=None, beamInDotHNormal=None):
        def for_one_polarization(polFactor):
            delta = np.sqrt((alpha**2 + polFactor**2 * chih * chih_ / b))
            if self.t is None:  # thick Bragg
                with np.errstate(divide='ignore', invalid='ignore'):
                    ra = chih * polFactor / (alpha+delta)
                ad = alpha - delta
                ad[ad == 0] = 1e-100
                rb = chih * polFactor / ad
                indB = np.where(np.isnan(ra))
                ra[indB] = rb[indB]
                indB = np.where(abs(rb) < abs(ra))
                ra[indB] = rb[indB]
                return ra / np.sqrt(abs(b))
            # finite thickness, crystal.py:598-616
            t = self.t * 1e7
            l = t * delta * k02 / 2. / kHs  # noqa: E741
            with np.errstate(over='ignore'):
                if self.geom.startswith('Bragg'):
                    if self.geom.endswith('transmitted'):
                        ra = 1 / (np.cos(l) - 1j * alpha * np.sin(l) / delta) * \
                            np.exp(1j * k02 * t * (chi0 - alpha*b) / 2 / k0s)
                    else:
                        ra = chih * polFactor / (alpha + 1j*delta / np.tan(l))
                else:  # Laue
                    if self.geom.endswith('transmitted'):
                        ra = (np.cos(l) + 1j * alpha * np.sin(l) / delta) * \
                            np.exp(1j * k02 * t * (chi0 - alpha*b) / 2 / k0s)
                    else:
                        ra = chih * polFactor * np.sin(l) / delta * np.exp(1j * k02 * t * (chi0 - alpha*b) / 2 / k0s)
            if not self.geom.endswith('transmitted'):
                ra /= np.sqrt(abs(b))
            return ra

        waveLength = CH / E
        k = PI2 / waveLength
        k0s = -beamInDotNormal * k
        if beamOutDotNormal is None:
            beamOutDotNormal = -beamInDotNormal
        kHs = -beamOutDotNormal * k
        if beamInDotHNormal is None:
            beamInDotHNormal = beamInDotNormal
        crystd = self.d
        HH = PI2 / crystd
        k0H = abs(beamInDotHNormal) * HH * k
        k02 = k**2
        H2 = HH**2
        kHs0 = kHs == 0
        kHs[kHs0] = 1
        b = k0s / kHs
        b[kHs0] = -1
        F0, Fhkl, Fhkl_, chi0, chih, chih_ = self.get_F_chi(E, 0.5/crystd)
        thetaB = self.get_Bragg_angle(E)
        alpha = (H2/2 - k0H) / k02 + chi0/2 * (1/b - 1)
        curveS = for_one_polarization(1.)
        polFactor = np.cos(2. * thetaB)
        curveP = for_one_polarization(polFactor)
        return curveS, curveP


class Multilayer(object):
    """Multilayer.get_amplitude, geom 'reflected', numpy branch (materials/multilayer.py:254-508; thickness laws
    set_dti / set_dbi :187-215).  tLayer, bLayer, substrate: oracle Material."""
    kind = 'multilayer'
    geom = 'reflected'

    def __init__(self, tLayer, tThickness, bLayer, bThickness, nPairs, substrate, tThicknessLow=0., bThicknessLow=0.,
                 idThickness=0., power=2., substRoughness=0., substThickness=np.inf, geom='reflected'):
        self.tLayer, self.bLayer, self.substrate = tLayer, bLayer, substrate
        self.substRoughness = substRoughness
        self.substThickness = substThickness
        self.geom = geom
        self.nPairs = int(nPairs)
        self.power = power
        self.tThicknessHigh, self.tThicknessLow = float(tThickness), float(tThicknessLow)
        self.bThicknessHigh, self.bThicknessLow = float(bThickness), float(bThicknessLow)
        self.idThickness = idThickness
        self.d = float(self.tThicknessHigh + self.bThicknessHigh)
        self.dti = self._graded(self.tThicknessHigh, self.tThicknessLow)
        self.dbi = self._graded(self.bThicknessHigh, self.bThicknessLow)

    def _graded(self, high, low):
        if low:
            layers = np.arange(1, self.nPairs+1)
            qRoot = (high/low)**(1./self.power)
            qB = (self.nPairs-qRoot) / (qRoot-1.)
            qA = high * (qB+1)**self.power
            return qA * (qB+layers)**(-self.power)
        return np.ones(self.nPairs) * float(high)

    def fresnel_terms(self, E, beamInDotNormal):
        """The interface amplitudes and layer wave vectors the reference hands to its accelerator kernel
        (multilayer.py:364-409, :471-477)."""
        k = E / CHBAR
        nt = self.tLayer.get_refractive_index(E).conjugate() if self.tLayer else 1.  # multilayer.py:373
        nb = self.bLayer.get_refractive_index(E).conjugate()
        ns = self.substrate.get_refractive_index(E).conjugate()
        Q = 2 * k * abs(beamInDotNormal)
        Q2 = Q**2
        k28 = 8 * k**2
        Qt = (Q2 + (nt-1)*k28)**0.5
        Qb = (Q2 + (nb-1)*k28)**0.5
        Qs = (Q2 + (ns-1)*k28)**0.5
        id2 = self.idThickness**2
        roughvt = np.exp(-0.5 * Q * Qt * id2)
        rvt_s = np.complex128((Q-Qt) / (Q+Qt) * roughvt)
        rvt_p = np.complex128((Q*nt - Qt/nt) / (Q*nt + Qt/nt) * roughvt)
        roughtb = np.exp(-0.5 * Qt * Qb * id2)
        rtb_s = np.complex128((Qt-Qb) / (Qt+Qb) * roughtb)
        rtb_p = np.complex128((Qt/nt*nb - Qb/nb*nt) / (Qt/nt*nb + Qb/nb*nt) * roughtb)
        rmsbs = id2 if self.tLayer else self.substRoughness**2  # multilayer.py:401
        roughbs = np.exp(-0.5 * Qb * Qs * rmsbs)
        rbs_s = np.complex128((Qb-Qs) / (Qb+Qs) * roughbs)
        rbs_p = np.complex128((Qb/nb*ns - Qs/ns*nb) / (Qb/nb*ns + Qs/ns*nb) * roughbs)
        out = dict(rbs_s=rbs_s, rbs_p=rbs_p, rtb_s=rtb_s, rtb_p=rtb_p, rvt_s=rvt_s, rvt_p=rvt_p, Qt=Qt, Qb=Qb, nt=nt)
        if 'tran' in self.geom:  # multilayer.py:388-417
            out.update(
                tvt_s=np.complex128(2*Q / (Q+Qt) * roughvt),
                tvt_p=np.complex128(2*Q*nt / (Q*nt + Qt/nt) * roughvt),
                ttb_s=np.complex128(2*Qt / (Qt+Qb) * roughtb),
                ttb_p=np.complex128(2*Qt/nt*nb / (Qt/nt*nb + Qb/nb*nt) * roughtb),
                tbt_s=np.complex128(2*Qb / (Qt+Qb) * roughtb),
                tbt_p=np.complex128(2*Qb/nb*nt / (Qt/nt*nb + Qb/nb*nt) * roughtb),
                tbs_s=np.complex128(2*Qb / (Qb+Qs) * roughbs),
                tbs_p=np.complex128(2*Qb/nb*ns / (Qb/nb*ns + Qs/ns*nb) * roughbs),
                rsv_s=np.complex128((Qs-Q) / (Qs+Q) * roughbs),
                rsv_p=np.complex128((Qs/ns - Q*ns) / (Qs/ns + Q*ns) * roughbs),
                tsv_s=np.complex128(2*Qs / (Qs+Q) * roughbs),
                tsv_p=np.complex128(2*Qs/ns / (Qs/ns + Q*ns) * roughbs), Qs=Qs)
        return out

    def recursion_tran(self, f):
        """The loop of Multilayer.get_amplitude for geom 'transmitted' (multilayer.py:419-462) == kernel
        get_amplitude_graded_multilayer_tran: the substrate is one more layer, then vacuum."""
        rj_s, rj_p = f['rsv_s'], f['rsv_p']
        tj_s, tj_p = f['tsv_s'], f['tsv_p']
        rbt_s, rbt_p = -f['rtb_s'], -f['rtb_p']
        ti_s = np.zeros_like(rj_s)
        ti_p = np.zeros_like(rj_p)
        for i in reversed(range(2*self.nPairs+1)):
            if i % 2 == 0:
                if i == 0:
                    rij_s, rij_p, tij_s, tij_p = f['rvt_s'], f['rvt_p'], f['tvt_s'], f['tvt_p']
                    iQT = f['Qt'] * self.dti[i//2]
                elif i == 2*self.nPairs:
                    rij_s, rij_p, tij_s, tij_p = f['rbs_s'], f['rbs_p'], f['tbs_s'], f['tbs_p']
                    iQT = f['Qs'] * self.substThickness
                else:
                    rij_s, rij_p, tij_s, tij_p = rbt_s, rbt_p, f['tbt_s'], f['tbt_p']
                    iQT = f['Qt'] * self.dti[i//2]
            else:
                rij_s, rij_p, tij_s, tij_p = f['rtb_s'], f['rtb_p'], f['ttb_s'], f['ttb_p']
                iQT = f['Qb'] * self.dbi[i//2]
            p1i = np.complex128(np.exp(0.5j*iQT))
            p2i = p1i**2
            rj2i_s = rj_s * p2i
            rj2i_p = rj_p * p2i
            ri_s = (rij_s + rj2i_s) / (1 + rij_s*rj2i_s)
            ri_p = (rij_p + rj2i_p) / (1 + rij_p*rj2i_p)
            ti_s = tij_s * tj_s * p1i / (1 + rij_s*rj2i_s)
            ti_p = tij_p * tj_p * p1i / (1 + rij_p*rj2i_p)
            tj_s, tj_p = ti_s, ti_p
            rj_s, rj_p = ri_s, ri_p
        return ti_s, ti_p

    def recursion(self, f):
        rj_s, rj_p = f['rbs_s'], f['rbs_p']
        rbt_s, rbt_p = -f['rtb_s'], -f['rtb_p']
        ri_s = np.zeros_like(rj_s)
        ri_p = np.zeros_like(rj_p)
        for i in reversed(range(2*self.nPairs)):
            if i % 2 == 0:
                if i == 0:
                    rij_s, rij_p = f['rvt_s'], f['rvt_p']
                else:
                    rij_s, rij_p = rbt_s, rbt_p
                iQT = f['Qt'] * self.dti[i//2]
            else:
                rij_s, rij_p = f['rtb_s'], f['rtb_p']
                iQT = f['Qb'] * self.dbi[i//2]
            p1i = np.complex128(np.exp(0.5j*iQT))
            p2i = p1i**2
            rj2i_s = rj_s * p2i
            rj2i_p = rj_p * p2i
            ri_s = (rij_s + rj2i_s) / (1 + rij_s*rj2i_s)
            ri_p = (rij_p + rj2i_p) / (1 + rij_p*rj2i_p)
            rj_s, rj_p = ri_s, ri_p
        return ri_s, ri_p

    def get_amplitude(self, E, beamInDotNormal, x=None, y=None):
        f = self.fresnel_terms(E, beamInDotNormal)
        if 'tran' in self.geom:
            return self.recursion_tran(f)
        ri_s, ri_p = self.recursion(f)
        nt = f['nt']
        nn = nt[0] if isinstance(nt, np.ndarray) else nt
        if (nn - 1) > 0:
            return ri_s.conjugate(), ri_p.conjugate()
        return ri_s, ri_p


class Coated(Multilayer):
    """Coated (materials/multilayer.py:528-584): a single reflective layer on a substrate = a Multilayer without top
    layer (vacuum, zero thickness), nPairs = 1, idThickness = the surface roughness; kind 'mirror', so OE.reflect
    treats it as a mirror (specular reflection; get_amplitude(E, beamInDotNormal, fromVacuum), reflect.py:1003-1025)."""
    kind = 'mirror'

    def __init__(self, coating, cThickness, substrate, surfaceRoughness=0., substRoughness=0.):
        Multilayer.__init__(self, None, 0., coating, cThickness, 1, substrate, idThickness=surfaceRoughness,
                            substRoughness=substRoughness)
