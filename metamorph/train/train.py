from metamorph_b200.train.train import DataArguments, ModelArguments, TrainingArguments, train  # noqa: F401

if __name__ == "__main__":
    train()
