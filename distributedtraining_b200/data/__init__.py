from .synthetic import ByteTokenizer, SyntheticMNIST, SyntheticTokens, WikitextDataset, custom_collate_fn, make_loader  # noqa: F401
