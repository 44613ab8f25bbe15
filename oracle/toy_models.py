"""Small models for the CPU test-suite and the golden-fixture generator."""

from workloads.unet_skeleton import TOY, UNetSkeleton


def ToyUNet():
    return UNetSkeleton(TOY)
