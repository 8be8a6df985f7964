from .synthetic import ByteTokenizer, SyntheticMNIST, SyntheticTokens, WikitextDataset, custom_collate_fn, make_loader  # noqa: F401
from .text import PinnedLoader, build_text_loader, build_tokenizer, read_lines  # noqa: F401
