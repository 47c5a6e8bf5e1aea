class GlobalHydra:
    _i = None
    @classmethod
    def instance(cls):
        cls._i = cls._i or cls(); return cls._i
    def is_initialized(self): return False
    def clear(self): pass
