# -*- coding: utf-8 -*-
"""Physical constants used by the hot path.

The *numerical values* must be the ones xrt computes with (xrt/backends/raycing/physconsts.py), because
quantities such as k = E / CHBAR * 1e7 enter phases of ~1e11 rad: tests/test_abi.py compares every value below
bit for bit with tests/golden/constants.npz (dumped from the reference)."""
import math as _math

PI = _math.pi                      # 3.141592653589793
PI2 = 2.0 * _math.pi               # 6.283185307179586
SQRT2PI = PI2 ** 0.5
SQ3 = 1.7320508075688772935274463415059

# CGS / SI mixture inherited from xrt (lengths in cm here, mm / Angstrom at the call sites)
C = 2.99792458e10                  # speed of light [cm/s]
SIE0 = 1.602176565e-19             # elementary charge [C]
EV2ERG = 1.602176565e-12           # 1 eV in erg
HPLANCK = 6.626069573e-27          # Planck constant [erg s]
M0 = 9.109383701528e-28            # electron mass [g]
SIM0 = 9.109383701528e-31          # electron mass [kg]
M0C2 = 0.510998928                 # electron rest energy [MeV]
R0 = 2.817940285e-5                # classical electron radius [Angstrom]
AVOGADRO = 6.02214199e23           # [1/mol]
FINE_STR = 1 / 137.03599976        # fine-structure constant

# derived exactly the way xrt derives them (operation order matters for the last bit)
E0 = SIE0 * C / 10                 # elementary charge [esu]
SIC = C * 1e-2                     # speed of light [m/s]
CHeVcm = HPLANCK * C / EV2ERG      # h c [eV cm]
CH = CHeVcm * 1e8                  # h c [eV Angstrom]  = 12398.419297617678
CHBAR = CH / PI2                   # hbar c [eV Angstrom] = 1973.2697177417986
K2B = 2 * PI * M0 * C**2 * 0.001 / E0   # undulator K -> B0[T] * period[mm]

# tabulated by xrt, not derived
E2W = 1519267514747457.9195337718065469   # eV -> rad/s
E2WC = 5067.7309392068091                 # eV -> 1/mm  (E2W / c[mm/s])
EMC = 0.5866791802416487                  # e / (m c) in xrt's trajectory units
