class ConfigStore:
    _i = None
    @classmethod
    def instance(cls):
        cls._i = cls._i or cls(); return cls._i
    def store(self, *a, **k): pass
