// xml_dom.h — a small in-memory XML element tree, enough for the diffusion-curve scene files.
// Fills the role rapidxml plays in the reference (optixHello.cpp:108-111, doc.parse<0>): elements and
// attributes only; DOCTYPE, comments, processing instructions and character data are skipped, exactly
// the nodes parse<0> does not create. Children and attributes keep document order; lookup by name is
// case-sensitive and returns the first match, like first_node(name)/first_attribute(name).
//
// Built for the 100 000-curve scene (76 MB of XML, 1.5 M elements, 4 M attributes): the document owns one
// copy of the text and parses it in place — names and values are NUL-terminated inside that buffer, entities are
// decoded where they stand (a decoded value is never longer than its source) — and elements and attributes
// come from bump arenas, so there is no allocation per node and the whole tree is freed in a handful of calls.
#ifndef RDC_XML_DOM_H
#define RDC_XML_DOM_H

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace rdc {

struct XmlError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

struct XmlAttr {
  const char* name;
  const char* value;
};

struct XmlElement {
  const char* name = "";
  const XmlAttr* attrs = nullptr;  // n_attrs of them, document order
  uint32_t n_attrs = 0;
  uint32_t n_children = 0;
  const XmlElement* first_child = nullptr;
  const XmlElement* next_sibling = nullptr;

  // value of the first attribute called `key`, or nullptr
  const char* attr(const char* key) const {
    for (uint32_t i = 0; i < n_attrs; ++i)
      if (std::strcmp(attrs[i].name, key) == 0) return attrs[i].value;
    return nullptr;
  }
  // first child element called `key`, or nullptr
  const XmlElement* child(const char* key) const {
    for (const XmlElement* c = first_child; c; c = c->next_sibling)
      if (std::strcmp(c->name, key) == 0) return c;
    return nullptr;
  }
};

// Owns the text and the tree. Elements and strings live as long as the document.
class XmlDocument {
 public:
  const XmlElement* root = nullptr;  // first top-level element

  XmlDocument() = default;
  XmlDocument(const XmlDocument&) = delete;
  XmlDocument& operator=(const XmlDocument&) = delete;

  std::vector<char> text;  // the document, modified in place while parsing
  void* allocate(size_t bytes);  // 8-byte aligned, zero-initialised, freed with the document

 private:
  std::vector<std::unique_ptr<uint64_t[]>> chunks_;
  size_t used_ = 0, capacity_ = 0;
};

// Parses a whole document held in memory (it is copied).
std::unique_ptr<XmlDocument> xml_parse(const char* text, size_t len);
// Reads the file and parses it. Throws XmlError when the file cannot be read or is malformed.
std::unique_ptr<XmlDocument> xml_parse_file(const std::string& path);
// Canonical dump (one line per element, attributes in document order) used by the parser tests.
void xml_dump(const XmlElement& e, int depth, std::string& out);

}  // namespace rdc

#endif
