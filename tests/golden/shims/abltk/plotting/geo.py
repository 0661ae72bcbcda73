import math

_R = 6371000.0


def latlon_to_xy(lat, lon, ref_lat, ref_lon):
    """Equirectangular stand-in; reproduces the reference goldens to 2.7e-12."""
    x = _R * math.radians(lon - ref_lon) * math.cos(math.radians(ref_lat))
    y = _R * math.radians(lat - ref_lat)
    return x, y
