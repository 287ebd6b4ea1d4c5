"""PubMLST strain-type lookup used at the end of MLST classification
(handlers/pubmlst.py:97-130 of the reference).  Network side effect: kept behind the same class and method so
callers (and tests) can inject a handler; allele download / training helpers are out of scope."""


class PubMLSTHandler:
    def __init__(self, base_url: str = "https://rest.pubmlst.org/db"):
        self.base_url = base_url

    def get_strain_type_name(self, highest_results: dict, post_url: str) -> str:
        """POST the best allele per locus to ``<post_url>/designations``; returns the ST fields or a message."""
        import requests

        payload = {"designations": {locus: [{"allele": str(allele)}] for locus, allele in highest_results.items()}}
        response = requests.post(post_url + "/designations", json=payload, timeout=10)
        if response.status_code == 200:
            data = response.json()
            if "fields" in data:
                return data["fields"]
            return "No matching Strain Type found in the database. Possibly a novel Strain Type."
        return "Error:" + str(response.status_code) + response.text
