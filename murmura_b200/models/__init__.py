"""Bundled model families + zero-argument factories usable from YAML (``model.factory``)."""
from murmura_b200.models.mlp import (MLP, EvidentialHead, EvidentialMLP, EvidentialHARClassifier,
                                     EvidentialPAMAP2Classifier, EvidentialPPGDaLiAClassifier,
                                     EvidentialLoss, compute_uncertainty)
from murmura_b200.models.cnn import (LEAFFEMNISTModel, LEAFCelebAModel, FEMNISTTiny, FEMNISTSmall,
                                     FEMNISTBaseline, FEMNISTLarge, FEMNISTXLarge, CIFARCNN,
                                     get_model_variant)
from murmura_b200.models.resnet import ResNet18, BasicBlock


def mlp(input_dim: int = 784, hidden_dims=(200,), num_classes: int = 10):
    return MLP(input_dim=input_dim, hidden_dims=tuple(hidden_dims), num_classes=num_classes)


def cifar_cnn(num_classes: int = 10):
    return CIFARCNN(num_classes=num_classes)


def resnet18(num_classes: int = 10):
    return ResNet18(num_classes=num_classes)


def femnist_cnn(num_classes: int = 62, variant: str = "baseline"):
    return get_model_variant(variant, num_classes=num_classes)


__all__ = ["MLP", "EvidentialHead", "EvidentialMLP", "EvidentialHARClassifier",
           "EvidentialPAMAP2Classifier", "EvidentialPPGDaLiAClassifier", "EvidentialLoss",
           "compute_uncertainty", "LEAFFEMNISTModel", "LEAFCelebAModel", "FEMNISTTiny", "FEMNISTSmall",
           "FEMNISTBaseline", "FEMNISTLarge", "FEMNISTXLarge", "CIFARCNN", "get_model_variant",
           "ResNet18", "BasicBlock", "mlp", "cifar_cnn", "resnet18", "femnist_cnn"]
