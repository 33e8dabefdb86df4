"""decord stand-in (video decode is off the path); util.py:24 calls decord.bridge.set_bridge at import."""


class _Bridge:
    @staticmethod
    def set_bridge(name):
        pass


bridge = _Bridge()
