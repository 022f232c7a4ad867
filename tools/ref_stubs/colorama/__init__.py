"""Import stub used only by tools/gen_golden*.py to import the reference in this container."""


class _C:
    def __getattr__(self, k):
        return ''


Fore = _C()
Back = _C()
Style = _C()


def init(*a, **k):
    pass
