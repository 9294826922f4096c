from murmura_b200.cli import app

if __name__ == "__main__":
    app()
