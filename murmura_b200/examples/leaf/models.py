"""FEMNIST model-size variants (re-exported from :mod:`murmura_b200.models.cnn`)."""
from murmura_b200.models.cnn import (FEMNISTBaseline, FEMNISTLarge, FEMNISTSmall, FEMNISTTiny, FEMNISTXLarge,
                                     get_model_variant)


def compare_model_sizes() -> None:
    for name in ("tiny", "small", "baseline", "large", "xlarge"):
        m = get_model_variant(name)
        print(f"{name:>9}: {m.parameter_count():,} parameters")


__all__ = ["FEMNISTTiny", "FEMNISTSmall", "FEMNISTBaseline", "FEMNISTLarge", "FEMNISTXLarge",
           "get_model_variant", "compare_model_sizes"]

if __name__ == "__main__":
    import torch
    compare_model_sizes()
    for name in ("tiny", "small", "baseline", "large", "xlarge"):
        assert get_model_variant(name)(torch.zeros(2, 1, 28, 28)).shape == (2, 62)
    print("forward shapes OK")
