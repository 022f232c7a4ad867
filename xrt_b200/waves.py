# -*- coding: utf-8 -*-
"""Wave propagation with xrt's call structure (xrt/backends/raycing/waves.py:505-831):

    wave = screen.prepare_wave(slit, xmesh, zmesh)      # receiving points      (screens.py:304-365)
    waveOnSlit = slit.prepare_wave(prevOE, nrays)       # diffracting samples   (apertures.py:467-499)
    glo = waves.diffract(oeLocal, wave)                 # N x M Kirchhoff sum + post-processing (waves.py:606-831)

The N x M sum runs in the hand-written CUDA kernel (csrc/kirchhoff.cu) on device-resident arrays; the O(M + N)
bookkeeping around it (accumulators, de-phasing of the direction vectors, flux normalisation, frame rotations)
is expressed with elementwise torch operations on the same device -- one rounding per operation, like the
numpy statements it follows.  The footprint area uses scipy's ConvexHull on the host exactly like the reference
(waves.py:649-670) when the local beam does not carry an `area` yet.

Supported on this path: diffracting element = RectangularAperture or an OE of the hot path (flat, toroid,
elliptical parametric: the traced local beam is the set of samples); receiving element = Screen or such an OE
(OE.prepare_wave / propagate_wave, oes/reflect.py:266-449, live in xrt_b200/oes.py).
"""
import numpy as np
import torch

from . import kernels
from .beams import Beam
from .physconsts import CHBAR, CH

F64 = torch.float64
C128 = torch.complex128


def _rotate_z(x, y, cosangle, sinangle):
    return cosangle*x - sinangle*y, sinangle*x + cosangle*y


def _refuse_two_surface(oe, what):
    """A DCM / Plate / lens as the diffracting or the previous element needs the second-surface frame of the
    reference (is2ndXtal: yaw = -pi, dt = -t, pitch + bragg, cryst2* angles; waves.py:545-566, oes/base.py:1169-1186),
    which is outside the hot path: refuse it instead of using the single-surface frame silently."""
    if hasattr(oe, 'cryst2pitch') or hasattr(oe, 'bragg') and hasattr(oe, 'cryst2perpTransl'):
        raise NotImplementedError('{0}: two-surface elements (DCM, Plate, lenses) as the diffracting / previous element '
                                  'are outside the hot path (second-surface frame of waves.py:545-566)'.format(what))


def prepare_wave(fromOE, wave, xglo, yglo, zglo):
    """waves.prepare_wave (waves.py:505-584): receiving points -> fromOE-local coordinates, accumulators."""
    n = len(wave)
    dev = wave.device
    z0 = lambda dt: torch.zeros(n, dtype=dt, device=dev)  # noqa: E731
    wave.Es, wave.Ep = z0(C128), z0(C128)
    for k in ('EsAcc', 'EpAcc', 'aEacc', 'bEacc', 'cEacc'):
        setattr(wave, k, z0(C128))
    wave.Jss.zero_()
    wave.Jpp.zero_()
    wave.Jsp.zero_()
    x = xglo.clone() - fromOE.center[0]
    y = yglo.clone() - fromOE.center[1]
    z = zglo.clone() - fromOE.center[2]
    a0, b0 = fromOE.bl.sinAzimuth, fromOE.bl.cosAzimuth
    x, y = _rotate_z(x, y, b0, a0)
    if hasattr(fromOE, 'rotationSequence'):  # OE
        from .oes import _rotation_list
        _refuse_two_surface(fromOE, 'prepare_wave')
        rots = _rotation_list(fromOE.rotationSequence, -fromOE.pitch, -(fromOE.roll+fromOE.positionRoll),
                              -fromOE.yaw)
        if fromOE.extraPitch or fromOE.extraRoll or fromOE.extraYaw:
            rots += _rotation_list(fromOE.extraRotationSequence, -fromOE.extraPitch, -fromOE.extraRoll,
                                   -fromOE.extraYaw)
        for ax, c, s in rots:
            if ax == 0:
                y, z = c*y - s*z, s*y + c*z
            elif ax == 1:
                x, z = c*x + s*z, -s*x + c*z
            else:
                x, y = c*x - s*y, s*x + c*y
    wave.xDiffr, wave.yDiffr, wave.zDiffr = x, y, z
    wave.rDiffr = (x**2 + y**2 + z**2)**0.5
    wave.a = x / wave.rDiffr
    wave.b = y / wave.rDiffr
    wave.c = z / wave.rDiffr
    wave.path.zero_()
    wave.fromOE = fromOE
    wave.beamReflRays = 0
    wave.beamReflSumJ = 0.
    wave.beamReflSumJnl = 0.
    wave.diffract_repeats = 0
    return wave


def screen_prepare_wave(screen, prevOE, dim1, dim2, dy=0, condition=None, device=None):
    """Screen.prepare_wave (screens.py:304-365) for a flat screen."""
    d1s, d2s = np.meshgrid(dim1, dim2)
    d1s = d1s.flatten()
    d2s = d2s.flatten()
    try:
        dS = (dim1[1] - dim1[0]) * (dim2[1] - dim2[0])
    except (IndexError, TypeError):
        dS = 1.
    if condition is not None:
        d1s, d2s = condition(d1s, d2s)
    nrays = len(d1s)
    xglo, yglo, zglo = screen.local_to_global(x=d1s, z=d2s)
    wave = Beam(nrays, withAmplitudes=True, device=device)
    wave.state.fill_(1)
    dev = wave.device
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)  # noqa: E731
    wave.x = t(d1s)
    wave.y = t(np.zeros_like(d1s) + dy)
    wave.z = t(d2s)
    wave.dS = dS
    wave.toOE = screen
    wave.area = (np.ones_like(d1s) * dS).sum()
    wave.parentId = getattr(prevOE, 'uuid', getattr(prevOE, 'name', None))
    return prepare_wave(prevOE, wave, t(xglo), t(yglo + dy), t(zglo))


def aperture_prepare_wave(slit, prevOE, nrays, xy=None, device=None, round_r=None):
    """RectangularAperture.prepare_wave (apertures.py:467-499) and, with *round_r*, RoundAperture.prepare_wave
    (apertures.py:848-875). *xy*: optional (nrays, 2) uniform deviates (the reference draws np.random.rand(nrays, 2));
    drawn with numpy's global stream when omitted."""
    nrays = int(nrays)
    if xy is None:
        xy = np.random.rand(nrays, 2)
    if round_r is not None:
        r = xy[:, 0]**0.5 * round_r
        phi = xy[:, 1] * 2*np.pi
        xl = r * np.cos(phi)
        zl = r * np.sin(phi)
        dX, dZ = np.pi * round_r**2, 1.  # wave.area = pi r^2
    else:
        dX = slit.limOptX[1] - slit.limOptX[0]
        dZ = slit.limOptY[1] - slit.limOptY[0]
        xl = xy[:, 0] * dX + slit.limOptX[0]
        zl = xy[:, 1] * dZ + slit.limOptY[0]
    wave = Beam(nrays, withAmplitudes=True, device=device)
    wave.state.fill_(1)
    dev = wave.device
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)  # noqa: E731
    wave.x = t(xl)
    wave.z = t(zl)
    wave.area = dX * dZ
    wave.dS = wave.area / nrays
    wave.toOE = slit
    wave.parentId = getattr(slit, 'uuid', slit.name)
    # local_to_global of the aperture (apertures.py:446-452) with y = 0
    xg = slit.center[0] + xl*slit.x[0] + 0.*slit.y[0] + zl*slit.z[0]
    yg = slit.center[1] + xl*slit.x[1] + 0.*slit.y[1] + zl*slit.z[1]
    zg = slit.center[2] + xl*slit.x[2] + 0.*slit.y[2] + zl*slit.z[2]
    return prepare_wave(prevOE, wave, t(xg), t(yg), t(zg))


def _hull_area(x, y):
    from scipy.spatial import ConvexHull
    pts = np.vstack((x, y)).T
    hull = ConvexHull(pts)
    outer = pts[hull.vertices, :]
    lines = np.hstack([outer, np.roll(outer, -1, axis=0)])
    return 0.5 * abs(sum(x1*y2-x2*y1 for x1, y1, x2, y2 in lines))


def _hull_vertices(x, y):
    from scipy.spatial import ConvexHull
    pts = np.vstack((x, y)).T
    if len(pts) < 3:
        return pts
    return pts[ConvexHull(pts).vertices, :]


def diffract(oeLocal, wave, targetOpenCL='CUDA', precisionOpenCL='float64', comm=None, samplesSharded=False):
    """waves.diffract (waves.py:606-831): fills *wave* with the diffracted field and returns the global beam.

    *comm* (xrt_b200.parallel.Comm, one process per GPU): *wave* holds THIS rank's shard of the receiving points --
    outputs are independent, every rank sums over all samples (the reference's run_parallel_max slices the
    receiving points the same way, myopencl.py:478-491, waves.py:864-894).  *samplesSharded* = True: *oeLocal* is this
    rank's ray shard of the diffracting samples too; their packed 128-byte records are all-gathered (one collective),
    the beam sums of waves.py:692-696 all-reduced and the footprint area taken as the hull of the ranks' hulls."""
    oe = wave.fromOE
    sharded = comm is not None and comm.world > 1 and samplesSharded
    dev = oeLocal.device
    good = oeLocal.state == 1
    goodlen = int(good.sum())
    goodlenAll = goodlen
    if sharded:
        goodlenAll = int(comm.allreduce_sum(torch.tensor([goodlen], dtype=torch.int64, device=dev)).item())
    if goodlenAll < 1e2 or len(wave.xDiffr) == 0:
        return Beam(goodlen, device=oeLocal.device)
    isOE = hasattr(oe, 'rotationSequence')
    if getattr(oeLocal, 'area', None) is None or oeLocal.area <= 0:
        second = oeLocal.y if isOE else oeLocal.z
        hx, hy = oeLocal.x[good].cpu().numpy(), second[good].cpu().numpy()
        if sharded:
            # the hull of the union is the hull of the ranks' hull vertices (a few hundred points per rank)
            import torch.distributed as dist
            mine = _hull_vertices(hx, hy)
            parts = [None] * comm.world
            dist.all_gather_object(parts, mine)
            allv = np.concatenate([p for p in parts if len(p)])
            oeLocal.area = _hull_area(allv[:, 0], allv[:, 1])
        else:
            oeLocal.area = _hull_area(hx, hy)
        if hasattr(oeLocal, 'areaFraction'):
            oeLocal.area *= oeLocal.areaFraction
    idx = torch.nonzero(good).flatten()
    sel = lambda t: t[idx].contiguous()  # noqa: E731
    x, y, z = sel(oeLocal.x), sel(oeLocal.y), sel(oeLocal.z)
    a, b, c = sel(oeLocal.a), sel(oeLocal.b), sel(oeLocal.c)
    if isOE:
        if getattr(oe, 'alpha', None):
            raise NotImplementedError('diffract from asymmetric crystals is the next row (SURVEY.md 8f)')
        if goodlen != len(oeLocal):
            # the reference's numpy integral broadcasts the un-masked normals against the masked samples
            # (waves.py:840): it only works when every ray of oeLocal is good -- same contract here
            raise ValueError('diffract from an OE needs a local beam whose rays are all good; filter it first')
        nX, nY, nZ = oe.local_n_device(oeLocal.x, oeLocal.y, oeLocal.z)  # waves.py:675-686
        nl = a*nX + b*nY + c*nZ
        nvec = None
    else:
        nvec = (0., 1., 0.)
        nl = a*nvec[0] + b*nvec[1] + c*nvec[2]
    Jtot = sel(oeLocal.Jss) + sel(oeLocal.Jpp)
    sums = torch.stack((Jtot.sum(), (Jtot * nl).sum()))
    if sharded:
        comm.allreduce_sum(sums)  # waves.py:692-696 over the whole beam
    wave.diffract_repeats += 1
    wave.beamReflRays += goodlenAll
    wave.beamReflSumJ += float(sums[0])
    wave.beamReflSumJnl += abs(float(sums[1]))

    E = sel(oeLocal.E)
    # waves.py:842.  torch divides a tensor by a Python scalar as a multiplication by its reciprocal, which is
    # one ulp off numpy's true division often enough to move k*r by 6e-5 rad: divide by a tensor instead
    k = torch.div(E, torch.full_like(E, CHBAR)) * 1e7
    pos = torch.stack((x, y, z, torch.zeros_like(x)), dim=1).contiguous()
    if nvec is not None:
        nrm = torch.tensor(nvec + (0.,), dtype=F64, device=x.device).repeat(goodlen, 1).contiguous()
    else:
        nrm = torch.stack((nX, nY, nZ, torch.zeros_like(nX)), dim=1).contiguous()
    if sharded:
        # every rank's own bound on |k r| (its samples x its receiving points), the largest decides for all
        pb = torch.tensor([kernels.phase_bound(wave.xDiffr, wave.yDiffr, wave.zDiffr, pos, k)], dtype=F64, device=dev)
        far = torch.stack((pos[:, :3].abs().max() if goodlen else torch.zeros((), dtype=F64, device=dev),
                           torch.stack((wave.xDiffr.abs().max(), wave.yDiffr.abs().max(), wave.zDiffr.abs().max())).max(),
                           k.abs().max() if goodlen else torch.zeros((), dtype=F64, device=dev)))
        comm.allreduce_max(far)  # samples of other ranks may lie farther: bound with the global extents
        pb = max(float(pb[0]), float(far[2] * (far[0] + far[1]) * 3**0.5))
        rec = kernels.kirchhoff_pack(nl, sel(oeLocal.Es), sel(oeLocal.Ep), k, pos, nrm)
        rec = comm.allgather_rows(rec)  # the single collective of the sharded Kirchhoff step: 128 B per sample
        out = kernels.kirchhoff_packed(wave.xDiffr.contiguous(), wave.yDiffr.contiguous(), wave.zDiffr.contiguous(),
                                       rec, convention='numpy', uniform_normal=nvec is not None,
                                       table_sincos=pb < kernels.PHASE_LIMIT_TABLE_SINCOS)
    else:
        out = kernels.kirchhoff(wave.xDiffr.contiguous(), wave.yDiffr.contiguous(), wave.zDiffr.contiguous(), nl,
                                sel(oeLocal.Es), sel(oeLocal.Ep), k, pos, nrm, convention='numpy',
                                uniform_normal=nvec is not None)
    Es, Ep, aE, bE, cE = out[0], out[1], out[2], out[3], out[4]

    wave.EsAcc += Es
    wave.EpAcc += Ep
    wave.aEacc += aE
    wave.bEacc += bE
    wave.cEacc += cE
    E0 = oeLocal.E[:1].clone() if len(oeLocal) else torch.zeros(1, dtype=F64, device=dev)
    if sharded:
        comm.broadcast(E0, 0)  # the first ray of the whole beam
    wave.E = torch.full_like(wave.E, float(E0[0]))
    wave.Es = wave.EsAcc.clone()
    wave.Ep = wave.EpAcc.clone()
    wave.Jss = (wave.Es * torch.conj(wave.Es)).real.clone()
    wave.Jpp = (wave.Ep * torch.conj(wave.Ep)).real.clone()
    wave.Jsp = wave.Es * torch.conj(wave.Ep)
    if isOE:
        # decided on the FIRST receiving point of the whole wave (waves.py:719-722): rank 0's when they are sharded
        useC = torch.tensor([1 if abs(complex(wave.cEacc[0])) > abs(complex(wave.bEacc[0])) else 0],
                            dtype=torch.int64, device=dev)
        if comm is not None and comm.world > 1:
            comm.broadcast(useC, 0)
        toRealComp = wave.cEacc if int(useC[0]) else wave.bEacc
    else:
        toRealComp = wave.bEacc
    toReal = torch.exp(-1j * torch.angle(toRealComp))
    wa = (wave.aEacc * toReal).real
    wb = (wave.bEacc * toReal).real
    wc = (wave.cEacc * toReal).real
    norm = (wa**2 + wb**2 + wc**2)**0.5
    norm = torch.where(norm == 0, torch.ones_like(norm), norm)
    wave.a, wave.b, wave.c = wa / norm, wb / norm, wc / norm

    fnorm = wave.dS * oeLocal.area * wave.beamReflSumJ
    de = wave.beamReflRays * wave.beamReflSumJnl * wave.diffract_repeats
    fnorm = fnorm / de if de > 0 else 0
    wave.Jss = wave.Jss * fnorm
    wave.Jpp = wave.Jpp * fnorm
    wave.Jsp = wave.Jsp * fnorm
    wave.Es = wave.Es * fnorm**0.5
    wave.Ep = wave.Ep * fnorm**0.5
    if hasattr(oeLocal, 'accepted'):
        wave.accepted = oeLocal.accepted
        wave.acceptedE = oeLocal.acceptedE
        wave.seeded = oeLocal.seeded
        lens = torch.tensor([len(wave), len(oeLocal)], dtype=torch.int64, device=dev)
        if comm is not None and comm.world > 1:
            own = lens.clone()
            comm.allreduce_sum(lens)
            if not sharded:
                lens[1] = own[1]
        wave.seededI = oeLocal.seededI * int(lens[0]) / int(lens[1])

    glo = Beam(copyFrom=wave)
    glo.parentId = getattr(oe, 'uuid', getattr(oe, 'name', None))
    glo.x, glo.y, glo.z = wave.xDiffr.clone(), wave.yDiffr.clone(), wave.zDiffr.clone()
    if isOE:
        oe.local_to_global(glo)  # oes/base.py:1165-1229
    elif hasattr(oe, 'local_to_global'):  # aperture: apertures.py:446-458
        gx = oe.center[0] + glo.x*oe.x[0] + glo.y*oe.y[0] + glo.z*oe.z[0]
        gy = oe.center[1] + glo.x*oe.x[1] + glo.y*oe.y[1] + glo.z*oe.z[1]
        gz = oe.center[2] + glo.x*oe.x[2] + glo.y*oe.y[2] + glo.z*oe.z[2]
        ga = glo.a*oe.x[0] + glo.b*oe.y[0] + glo.c*oe.z[0]
        gb = glo.a*oe.x[1] + glo.b*oe.y[1] + glo.c*oe.z[1]
        gc = glo.a*oe.x[2] + glo.b*oe.y[2] + glo.c*oe.z[2]
        glo.x, glo.y, glo.z, glo.a, glo.b, glo.c = gx, gy, gz, ga, gb, gc
    toOE = getattr(wave, 'toOE', None)
    if toOE is not None:
        # rotate abc into the local system of the receiving surface (waves.py:774-824); a Screen has no
        # rotationSequence, so only the azimuth rotation applies
        wave.a, wave.b, wave.c = glo.a.clone(), glo.b.clone(), glo.c.clone()
        wave.Jss, wave.Jpp, wave.Jsp = glo.Jss.clone(), glo.Jpp.clone(), glo.Jsp.clone()
        wave.Es, wave.Ep = glo.Es.clone(), glo.Ep.clone()
        a0, b0 = toOE.bl.sinAzimuth, toOE.bl.cosAzimuth
        wave.a, wave.b = _rotate_z(wave.a, wave.b, b0, a0)
        if hasattr(toOE, 'rotationSequence'):  # receiving surface is an OE (waves.py:786-824)
            from .oes import _rotation_list
            nX, nY, nZ = toOE.local_n_device(wave.x, wave.y, wave.z)
            rollAngle = (toOE.roll + toOE.positionRoll) + torch.atan2(nX, nZ)
            c, s = torch.cos(-rollAngle), torch.sin(-rollAngle)  # rotate_coherency_matrix(wave, ., -rollAngle)
            c2, s2, cs = c**2, s**2, c * s
            Jss, Jpp, Jsp = wave.Jss, wave.Jpp, wave.Jsp
            JssN = Jss*c2 + Jpp*s2 + 2*Jsp.real*cs
            JppN = Jss*s2 + Jpp*c2 - 2*Jsp.real*cs
            JspN = torch.complex((Jpp-Jss)*cs + Jsp.real*(c2-s2), Jsp.imag.clone())
            wave.Jss, wave.Jpp, wave.Jsp = JssN, JppN, JspN
            cosY, sinY = torch.cos(rollAngle), torch.sin(rollAngle)
            Es, Ep = wave.Es, wave.Ep  # rotate_y(Es, Ep, cosY, -sinY)
            wave.Es, wave.Ep = cosY*Es + (-sinY)*Ep, -(-sinY)*Es + cosY*Ep
            rots = _rotation_list(toOE.rotationSequence, -toOE.pitch, -toOE.roll-toOE.positionRoll, -toOE.yaw)
            if toOE.extraPitch or toOE.extraRoll or toOE.extraYaw:
                rots += _rotation_list(toOE.extraRotationSequence, -toOE.extraPitch, -toOE.extraRoll,
                                       -toOE.extraYaw)
            wa, wb, wc = wave.a, wave.b, wave.c
            for ax, cA, sA in rots:
                if ax == 0:
                    wb, wc = cA*wb - sA*wc, sA*wb + cA*wc
                elif ax == 1:
                    wa, wc = cA*wa + sA*wc, -sA*wa + cA*wc
                else:
                    wa, wb = cA*wa - sA*wb, sA*wa + cA*wb
            wave.a, wave.b, wave.c = wa, wb, wc
            norm = torch.abs(-wave.a*nX - wave.b*nY - wave.c*nZ)
            for b in (wave, glo):
                b.Jss = b.Jss * norm
                b.Jpp = b.Jpp * norm
                b.Jsp = b.Jsp * norm
                b.Es = b.Es * norm**0.5
                b.Ep = b.Ep * norm**0.5
    glo.createdByDiffract = True
    return glo
