#include "xml_dom.h"

#include <cstdio>
#include <cstdlib>

namespace rdc {

void* XmlDocument::allocate(size_t bytes) {
  const size_t words = (bytes + 7) / 8;
  if (used_ + words > capacity_) {
    const size_t chunk = words > (1u << 16) ? words : (1u << 16);  // 512 KB at a time
    chunks_.emplace_back(new uint64_t[chunk]());
    used_ = 0;
    capacity_ = chunk;
  }
  void* p = chunks_.back().get() + used_;
  used_ += words;
  return p;
}

namespace {

struct Parser {
  XmlDocument& doc;
  char* begin;
  char* p;
  char* end;  // *end == '\0' (the buffer has a terminator), so p[0] may always be read
  std::vector<XmlAttr> scratch;  // attributes of the tag being parsed
  int depth = 0;

  [[noreturn]] void fail(const std::string& what) const {
    size_t line = 1;
    for (const char* q = begin; q < p && q < end; ++q) line += *q == '\n';
    throw XmlError("xml: " + what + " at line " + std::to_string(line));
  }

  static bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }
  static bool is_name_char(char c) { return !(is_space(c) || c == '/' || c == '>' || c == '=' || c == '<' || c == '\0'); }

  void skip_space() {
    while (is_space(*p)) ++p;
  }
  bool starts_with(const char* s, size_t n) const { return (size_t)(end - p) >= n && std::memcmp(p, s, n) == 0; }

  void skip_until(const char* terminator) {
    const size_t n = std::strlen(terminator);
    while (p < end && !starts_with(terminator, n)) ++p;
    if (p >= end) fail(std::string("unterminated construct, expected ") + terminator);
    p += n;
  }

  // <!DOCTYPE ...> may nest an internal subset in [ ].
  void skip_doctype() {
    int nest = 0;
    while (p < end) {
      const char ch = *p;
      if (ch == '[') ++nest;
      else if (ch == ']') --nest;
      else if (ch == '>' && nest <= 0) {
        ++p;
        return;
      }
      ++p;
    }
    fail("unterminated DOCTYPE");
  }

  static char* append_utf8(char* out, unsigned long cp) {
    if (cp < 0x80) *out++ = (char)cp;
    else if (cp < 0x800) {
      *out++ = (char)(0xC0 | (cp >> 6));
      *out++ = (char)(0x80 | (cp & 0x3F));
    } else if (cp < 0x10000) {
      *out++ = (char)(0xE0 | (cp >> 12));
      *out++ = (char)(0x80 | ((cp >> 6) & 0x3F));
      *out++ = (char)(0x80 | (cp & 0x3F));
    } else {
      *out++ = (char)(0xF0 | (cp >> 18));
      *out++ = (char)(0x80 | ((cp >> 12) & 0x3F));
      *out++ = (char)(0x80 | ((cp >> 6) & 0x3F));
      *out++ = (char)(0x80 | (cp & 0x3F));
    }
    return out;
  }

  // Decodes the entities of [b, e) where they stand; returns the new end. A reference is at least as long as
  // what it stands for ("&#9;" -> 1 byte, "&#x10FFFF;" -> 4), except numeric references padded beyond reason
  // or out of range, which are kept verbatim like unknown names.
  static char* decode_entities(char* b, char* e) {
    char* out = b;
    while (b < e) {
      if (*b != '&') {
        *out++ = *b++;
        continue;
      }
      char* semi = (char*)std::memchr(b, ';', e - b);
      if (!semi) {
        *out++ = *b++;
        continue;
      }
      const size_t n = semi - (b + 1);
      const char* ent = b + 1;
      auto is = [&](const char* s) { return n == std::strlen(s) && std::memcmp(ent, s, n) == 0; };
      if (is("amp")) *out++ = '&';
      else if (is("lt")) *out++ = '<';
      else if (is("gt")) *out++ = '>';
      else if (is("quot")) *out++ = '"';
      else if (is("apos")) *out++ = '\'';
      else if (n > 1 && ent[0] == '#') {
        const std::string digits(ent + 1, n - 1);
        const unsigned long cp = (digits[0] == 'x' || digits[0] == 'X') ? std::strtoul(digits.c_str() + 1, nullptr, 16)
                                                                        : std::strtoul(digits.c_str(), nullptr, 10);
        if (cp <= 0x10FFFF) out = append_utf8(out, cp);  // at most 4 bytes; the reference is at least 4 ("&#9;")
        else {
          std::memmove(out, b, semi + 1 - b);
          out += semi + 1 - b;
        }
      } else {  // unknown entity: keep verbatim
        std::memmove(out, b, semi + 1 - b);
        out += semi + 1 - b;
      }
      b = semi + 1;
    }
    return out;
  }

  char* parse_name() {
    char* b = p;
    while (is_name_char(*p)) ++p;
    if (p == b) fail("expected a name");
    return b;
  }

  // Skips everything that is not an element start. Returns false at end of input or at a closing tag.
  bool seek_element() {
    for (;;) {
      char* lt = (char*)std::memchr(p, '<', end - p);  // character data
      if (!lt) {
        p = end;
        return false;
      }
      p = lt;
      if (starts_with("<!--", 4)) skip_until("-->");
      else if (starts_with("<![CDATA[", 9)) skip_until("]]>");
      else if (starts_with("<!", 2)) skip_doctype();
      else if (starts_with("<?", 2)) skip_until("?>");
      else if (starts_with("</", 2)) return false;
      else return true;
    }
  }

  XmlElement* parse_element() {
    if (++depth > 512) fail("elements nested deeper than 512");
    ++p;  // '<'
    XmlElement* el = new (doc.allocate(sizeof(XmlElement))) XmlElement();
    char* name = parse_name();
    char* name_end = p;
    const size_t first = scratch.size();
    bool empty = false;
    // Ends of names and values are recorded and only turned into terminators once the tag has been read: the
    // byte after a name may be the '/' or '>' that still has to be looked at.
    std::vector<char*>& ends = ends_;
    const size_t first_end = ends.size();
    for (;;) {
      skip_space();
      if (p >= end) {
        *name_end = '\0';
        fail(std::string("unterminated tag <") + name);
      }
      if (*p == '/') {
        ++p;
        if (*p != '>') fail("expected '>' after '/'");
        ++p;
        empty = true;
        break;
      }
      if (*p == '>') {
        ++p;
        break;
      }
      char* key = parse_name();
      char* key_end = p;
      skip_space();
      if (*p != '=') fail("expected '=' after attribute " + std::string(key, key_end));
      ++p;
      skip_space();
      const char quote = *p;
      if (quote != '"' && quote != '\'') fail("expected a quoted value for attribute " + std::string(key, key_end));
      ++p;
      char* value = p;
      char* close = (char*)std::memchr(p, quote, end - p);
      if (!close) fail("unterminated value of attribute " + std::string(key, key_end));
      char* value_end = std::memchr(value, '&', close - value) ? decode_entities(value, close) : close;
      p = close + 1;
      ends.push_back(key_end);
      ends.push_back(value_end);
      scratch.push_back(XmlAttr{key, value});
    }
    *name_end = '\0';
    for (size_t i = first_end; i < ends.size(); ++i) *ends[i] = '\0';
    ends.resize(first_end);
    el->name = name;
    el->n_attrs = (uint32_t)(scratch.size() - first);
    if (el->n_attrs) {
      XmlAttr* a = static_cast<XmlAttr*>(doc.allocate(sizeof(XmlAttr) * el->n_attrs));
      std::memcpy(a, scratch.data() + first, sizeof(XmlAttr) * el->n_attrs);
      el->attrs = a;
    }
    scratch.resize(first);
    if (!empty) {
      XmlElement* last = nullptr;
      while (seek_element()) {
        XmlElement* c = parse_element();
        if (last) last->next_sibling = c;
        else el->first_child = c;
        last = c;
        el->n_children++;
      }
      if (p >= end) fail(std::string("missing </") + el->name + ">");
      p += 2;  // "</"
      char* closing = parse_name();
      if ((size_t)(p - closing) != std::strlen(el->name) || std::memcmp(closing, el->name, p - closing) != 0)
        fail("</" + std::string(closing, p) + "> closes <" + el->name + ">");
      skip_space();
      if (*p != '>') fail("expected '>' in closing tag");
      ++p;
    }
    --depth;
    return el;
  }

  std::vector<char*> ends_;
};

}  // namespace

namespace {

// doc->text holds the document followed by one '\0'
void parse_in_place(XmlDocument& doc) {
  char* begin = doc.text.data();
  // an embedded NUL would end names early: treat it as the end of the document, like a C string
  const size_t visible = std::strlen(begin);
  Parser ps{doc, begin, begin, begin + visible, {}};
  if (visible >= 3 && (unsigned char)ps.p[0] == 0xEF && (unsigned char)ps.p[1] == 0xBB && (unsigned char)ps.p[2] == 0xBF)
    ps.p += 3;  // UTF-8 byte order mark
  if (!ps.seek_element()) throw XmlError("xml: document has no root element");
  doc.root = ps.parse_element();
}

}  // namespace

std::unique_ptr<XmlDocument> xml_parse(const char* text, size_t len) {
  auto doc = std::make_unique<XmlDocument>();
  doc->text.resize(len + 1);
  if (len) std::memcpy(doc->text.data(), text, len);
  doc->text[len] = '\0';
  parse_in_place(*doc);
  return doc;
}

std::unique_ptr<XmlDocument> xml_parse_file(const std::string& path) {
  std::FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) throw XmlError("xml: cannot open " + path);
  auto doc = std::make_unique<XmlDocument>();
  // straight into the document's buffer: the size when the file can tell it, growing reads otherwise (pipes)
  long size = -1;
  if (std::fseek(f, 0, SEEK_END) == 0) {
    size = std::ftell(f);
    std::rewind(f);
  }
  size_t have = 0;
  doc->text.resize(size > 0 ? (size_t)size + 1 : (1u << 16));
  for (;;) {
    if (have + 1 >= doc->text.size()) doc->text.resize(doc->text.size() * 2);
    const size_t n = std::fread(doc->text.data() + have, 1, doc->text.size() - 1 - have, f);
    if (n == 0) break;
    have += n;
  }
  std::fclose(f);
  doc->text.resize(have + 1);
  doc->text[have] = '\0';
  parse_in_place(*doc);
  return doc;
}

void xml_dump(const XmlElement& e, int depth, std::string& out) {
  out.append((size_t)depth, ' ');
  out += e.name;
  for (uint32_t i = 0; i < e.n_attrs; ++i) {
    out += ' ';
    out += e.attrs[i].name;
    out += '=';
    out += e.attrs[i].value;
  }
  out += '\n';
  for (const XmlElement* c = e.first_child; c; c = c->next_sibling) xml_dump(*c, depth + 1, out);
}

}  // namespace rdc
