"""Model lookup by slug (model_management.py:10-204 of the reference): same slugs, same JSON metadata."""

import re
import unicodedata
from json import dumps, loads
from pathlib import Path

from .definitions import get_xspect_model_path


def slugify(text: str) -> str:
    """python-slugify's default behaviour restated (the package is not a dependency here): quotes dropped,
    NFKD-folded to ASCII, lower-cased, digit-group commas removed, every other run of characters outside
    ``[-a-z0-9]`` becomes one '-', leading/trailing '-' stripped."""
    text = str(text)
    text = re.sub(r"[']+", "-", text)
    text = unicodedata.normalize("NFKD", text)
    text = "".join(c for c in text if not unicodedata.combining(c))
    text = text.encode("ascii", "ignore").decode("ascii").lower()
    text = re.sub(r"[']+", "", text)
    text = re.sub(r"(?<=\d),(?=\d)", "", text)
    text = re.sub(r"[^-a-z0-9]+", "-", text)
    return re.sub(r"-{2,}", "-", text).strip("-")


def get_genus_model_path(genus) -> Path:
    return get_xspect_model_path() / (slugify(genus) + "-genus.json")


def get_species_model_path(genus) -> Path:
    return get_xspect_model_path() / (slugify(genus) + "-species.json")


def get_mlst_model_path(organism: str, scheme: str) -> Path:
    return get_xspect_model_path() / (slugify(organism + "-" + scheme + "-mlst") + ".json")


def get_model_metadata(model: str | Path) -> dict:
    if isinstance(model, str):
        model_path = get_xspect_model_path() / (slugify(model) + ".json")
    elif isinstance(model, Path):
        model_path = model
    else:
        raise ValueError("Model must be a string (slug) or a Path object.")
    if not model_path.exists() or not model_path.is_file():
        raise ValueError(f"Model at {model_path} does not exist.")
    return loads(model_path.read_text(encoding="utf-8"))


def is_svm_model(model_slug: str) -> bool:
    return get_model_metadata(model_slug).get("model_class") == "ProbabilisticFilterSVMModel"


def _rewrite(model_slug: str, metadata: dict) -> None:
    (get_xspect_model_path() / (model_slug + ".json")).write_text(dumps(metadata, indent=4), encoding="utf-8")


def update_model_metadata(model_slug: str, author: str, author_email: str) -> None:
    metadata = get_model_metadata(model_slug)
    metadata["author"] = author
    metadata["author_email"] = author_email
    _rewrite(model_slug, metadata)


def update_model_display_name(model_slug: str, filter_id: str, display_name: str) -> None:
    metadata = get_model_metadata(model_slug)
    metadata["display_names"][filter_id] = display_name
    _rewrite(model_slug, metadata)


def get_models() -> dict[str, list[dict]]:
    by_type: dict[str, list] = {}
    for model_file in get_xspect_model_path().glob("*.json"):
        metadata = get_model_metadata(model_file)
        by_type.setdefault(metadata["model_type"], []).append(metadata["model_display_name"])
    return by_type


def get_model_display_names(model_slug: str) -> list[str]:
    return list(get_model_metadata(model_slug)["display_names"].values())


def get_available_mlst_schemes() -> dict[str, list[str]]:
    schemes: dict[str, list[str]] = {}
    for model_file in get_xspect_model_path().glob("*-mlst.json"):
        metadata = get_model_metadata(model_file)
        organism, scheme = metadata.get("organism"), metadata.get("model_display_name")
        if organism and scheme:
            schemes.setdefault(organism, []).append(scheme)
    return schemes
