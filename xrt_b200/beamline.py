# -*- coding: utf-8 -*-
"""Minimal `BeamLine` container: what the hot path reads from xrt's BeamLine
(xrt/backends/raycing/beamline.py:407-545): azimuth (sin/cos), height, the ordered lists of elements that give
every element its ordinal number (and thereby its `lostNum`, beamline.py:80-100), and `alignE`.
Flow bookkeeping, auto-alignment from beams, layouts and glow stay with xrt (out of scope)."""
import numpy as np


class BeamLine(object):
    def __init__(self, azimuth=0., height=0., alignE='auto'):
        self.azimuth = azimuth
        self.height = height
        self.alignE = alignE
        self.sources = []
        self.oes = []
        self.slits = []
        self.screens = []
        self.alarms = []
        self.name = ''
        self.oesDict = {}
        self.flowSource = 'legacy'

    @property
    def azimuth(self):
        return self._azimuth

    @azimuth.setter
    def azimuth(self, value):
        self._azimuth = value
        self.sinAzimuth = np.sin(value)  # beamline.py:495-498
        self.cosAzimuth = np.cos(value)
