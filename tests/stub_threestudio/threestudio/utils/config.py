"""utils/config.py:137-139 of the reference: `parse_structured(fields, cfg) = OmegaConf.structured(fields(**cfg))`.
Without omegaconf the dataclass instance itself is the structured config; an unknown key still raises (TypeError
instead of omegaconf's ConfigKeyError)."""


def parse_structured(fields, cfg=None):
    return fields(**dict(cfg or {}))
