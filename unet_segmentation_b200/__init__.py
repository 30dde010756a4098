"""B200-native (sm_100a) U-Net training / inference hot path behind the reference's
`UNet` / `WeightedCrossEntropyLoss` interface. See DESIGN.md."""
__version__ = "0.1.0"
