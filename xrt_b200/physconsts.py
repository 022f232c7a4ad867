# -*- coding: utf-8 -*-
"""Physical constants of the raycing hot path. The numerical values are the ones xrt uses
(xrt/backends/raycing/physconsts.py:5-42) so that k = E/CHBAR*1e7 etc. are bit-identical; the tests
compare them against tests/golden/constants.npz dumped from the reference."""
PI = 3.1415926535897932384626433832795
PI2 = 6.283185307179586476925286766559
SQRT2PI = PI2**0.5
SIE0 = 1.602176565e-19
C = 2.99792458e10  # [cm/sec]
E0 = SIE0 * C / 10
M0 = 9.109383701528e-28  # [g]
SIM0 = 9.109383701528e-31
M0C2 = 0.510998928  # MeV
HPLANCK = 6.626069573e-27  # [erg*sec]
EV2ERG = 1.602176565e-12
K2B = 2 * PI * M0 * C**2 * 0.001 / E0
EMC = 0.5866791802416487
SIC = C * 1e-2
FINE_STR = 1 / 137.03599976
E2W = 1519267514747457.9195337718065469
E2WC = 5067.7309392068091
R0 = 2.817940285e-5  # A
AVOGADRO = 6.02214199e23  # atoms/mol
CHeVcm = HPLANCK * C / EV2ERG  # c*h [eV*cm]
CH = CHeVcm * 1e8  # c*h [eV*A]
CHBAR = CH / PI2  # c*h/(2pi) [eV*A]
