"""The MNIST ResNet-GP experiment scripts of the reference (exp_mnist_resnet/), on the B200
kernels: save_kernel (Gram blocks -> store), merge_h5_files (NaN-aware merge), classify_gp
(fp64 SPD solve + prediction on the GPU), and run (one process per GPU, torchrun)."""
