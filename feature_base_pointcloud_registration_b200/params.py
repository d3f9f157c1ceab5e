"""Reader for the reference's flat ``config/params.yaml`` (same knob names; reference
``include/utility.h:146-212`` holds the code defaults, ``config/params.yaml:19-67`` the deployed values)."""
import re

# code defaults of ParamServer (utility.h:164-198)
DEFAULTS = dict(N_SCAN=16, Horizon_SCAN=1800, edgeThreshold=0.1, surfThreshold=0.1,
                edgeFeatureMinValidNum=10, surfFeatureMinValidNum=100,
                odometrySurfLeafSize=0.2, mappingCornerLeafSize=0.2, mappingSurfLeafSize=0.2,
                z_tollerance=3.4028234663852886e38, rotation_tollerance=3.4028234663852886e38,
                numberOfCores=2, surroundingKeyframeSearchRadius=50.0)


def load_params_yaml(path):
    """Parse `key: value  # comment` lines; unknown keys are ignored, missing ones take the code defaults."""
    out = dict(DEFAULTS)
    try:
        import yaml
        with open(path) as f:
            data = yaml.safe_load(f) or {}
    except ImportError:  # tiny fallback reader for flat files
        data = {}
        with open(path) as f:
            for line in f:
                m = re.match(r"^\s*([A-Za-z_]\w*)\s*:\s*([^#]+?)\s*(#.*)?$", line)
                if m:
                    data[m.group(1)] = m.group(2)
    for k, dflt in DEFAULTS.items():
        if k in data:
            out[k] = type(dflt)(data[k])
    return out
