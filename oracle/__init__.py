"""TEST INFRASTRUCTURE ONLY.  CPU/PyTorch-fp32 restatement of the reference's cUNet generator,
discriminator and train step.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package; the product (weather-unet_b200/) never does."""
