"""The bindings of the reference's example configs, restated as data (the `.gin` files live in the read-only reference
tree, which is absent on the GPU box).  Citations: example_configs/<name>.gin of google/compare_gan @19922d3.
`tests/test_host_logic.py` checks, when the reference tree is present, that parsing these strings and parsing the
original files produce the same bindings."""

RESNET_CIFAR10 = """
dataset.name = "cifar10"
options.architecture = "resnet_cifar_arch"
options.batch_size = 64
options.gan_class = @ModularGAN
options.lamba = 1
options.training_steps = 40000
options.z_dim = 128
G.batch_norm_fn = @batch_norm
standardize_batch.decay = 0.9
standardize_batch.epsilon = 1e-5
options.disc_iters = 5
D.spectral_norm = True
loss.fn = @non_saturating
penalty.fn = @no_penalty
ModularGAN.g_lr = 0.0002
ModularGAN.g_optimizer_fn = @tf.train.AdamOptimizer
tf.train.AdamOptimizer.beta1 = 0.5
tf.train.AdamOptimizer.beta2 = 0.999
"""

SNDCGAN_CELEBAHQ128 = """
dataset.name = "celeb_a_hq_128"
options.architecture = "sndcgan_arch"
options.batch_size = 64
options.gan_class = @ModularGAN
options.lamba = 1
options.training_steps = 100000
options.z_dim = 128
G.batch_norm_fn = @batch_norm
standardize_batch.decay = 0.9
standardize_batch.epsilon = 1e-5
options.disc_iters = 1
D.spectral_norm = True
loss.fn = @non_saturating
penalty.fn = @no_penalty
ModularGAN.g_lr = 0.0002
ModularGAN.g_optimizer_fn = @tf.train.AdamOptimizer
tf.train.AdamOptimizer.beta1 = 0.5
tf.train.AdamOptimizer.beta2 = 0.999
"""

RESNET_LSUN_BEDROOM128 = """
dataset.name = "lsun-bedroom"
options.architecture = "resnet5_arch"
options.batch_size = 64
options.gan_class = @ModularGAN
options.lamba = 10
options.training_steps = 40000
options.z_dim = 128
G.batch_norm_fn = @batch_norm
standardize_batch.decay = 0.9
standardize_batch.epsilon = 1e-5
options.disc_iters = 5
D.spectral_norm = False
loss.fn = @wasserstein
penalty.fn = @wgangp_penalty
ModularGAN.g_lr = 0.0001
ModularGAN.g_optimizer_fn = @tf.train.AdamOptimizer
tf.train.AdamOptimizer.beta1 = 0.5
tf.train.AdamOptimizer.beta2 = 0.9
"""

BIGGAN_IMAGENET128 = """
dataset.name = "imagenet_128"
options.z_dim = 120
options.architecture = "resnet_biggan_arch"
ModularGAN.conditional = True
options.batch_size = 2048
options.gan_class = @ModularGAN
options.lamba = 1
options.training_steps = 250000
weights.initializer = "orthogonal"
spectral_norm.singular_value = "auto"
G.batch_norm_fn = @conditional_batch_norm
G.spectral_norm = True
ModularGAN.g_use_ema = True
resnet_biggan.Generator.hierarchical_z = True
resnet_biggan.Generator.embed_y = True
standardize_batch.decay = 0.9
standardize_batch.epsilon = 1e-5
standardize_batch.use_moving_averages = False
options.disc_iters = 2
D.spectral_norm = True
resnet_biggan.Discriminator.project_y = True
loss.fn = @hinge
penalty.fn = @no_penalty
ModularGAN.g_lr = 0.0001
ModularGAN.g_optimizer_fn = @tf.train.AdamOptimizer
ModularGAN.d_lr = 0.0005
ModularGAN.d_optimizer_fn = @tf.train.AdamOptimizer
tf.train.AdamOptimizer.beta1 = 0.0
tf.train.AdamOptimizer.beta2 = 0.999
z.distribution_fn = @tf.random.normal
eval_z.distribution_fn = @tf.random.normal
run_config.iterations_per_loop = 500
run_config.save_checkpoints_steps = 2500
"""

DCGAN_CELEBA64 = """
dataset.name = "celeb_a"
options.architecture = "dcgan_arch"
options.batch_size = 64
options.gan_class = @ModularGAN
options.lamba = 1
options.training_steps = 100000
options.z_dim = 128
G.batch_norm_fn = @batch_norm
standardize_batch.decay = 0.9
standardize_batch.epsilon = 1e-5
options.disc_iters = 1
D.spectral_norm = False
loss.fn = @non_saturating
penalty.fn = @no_penalty
ModularGAN.g_lr = 0.0002
ModularGAN.g_optimizer_fn = @tf.train.AdamOptimizer
tf.train.AdamOptimizer.beta1 = 0.5
tf.train.AdamOptimizer.beta2 = 0.999
"""

CONFIGS = {
    "dcgan_celeba64": DCGAN_CELEBA64,
    "resnet_cifar10": RESNET_CIFAR10,
    "sndcgan_celebahq128": SNDCGAN_CELEBAHQ128,
    "resnet_lsun-bedroom128": RESNET_LSUN_BEDROOM128,
    "biggan_imagenet128": BIGGAN_IMAGENET128,
}
