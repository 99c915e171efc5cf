from metamorph_b200.train.train import train

if __name__ == "__main__":
    train()
