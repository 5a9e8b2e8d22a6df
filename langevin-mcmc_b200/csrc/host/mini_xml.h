// mini_xml.h -- a ~100-line XML reader for the Mitsuba-0.5 scene subset the reference parses
// with pugixml (src/parsescene.cpp).  Elements, attributes, nesting, comments, <?xml ...?>.
#pragma once
#include <string>
#include <vector>
#include <map>
#include <memory>
#include <stdexcept>
#include <cctype>

namespace lmc_host {

struct XmlNode {
    std::string name;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::vector<std::shared_ptr<XmlNode>> children;
    bool has(const std::string &k) const {
        for (auto &a : attrs) if (a.first == k) return true;
        return false;
    }
    std::string attr(const std::string &k) const {
        for (auto &a : attrs) if (a.first == k) return a.second;
        return std::string();
    }
};

class XmlParser {
  public:
    explicit XmlParser(const std::string &text) : s(text), p(0) {}
    std::shared_ptr<XmlNode> parse() {
        auto root = std::make_shared<XmlNode>();
        root->name = "#document";
        parseChildren(*root, "");
        return root;
    }

  private:
    const std::string &s;
    size_t p;
    void skipWs() { while (p < s.size() && isspace((unsigned char)s[p])) p++; }
    bool startsWith(const char *t) const { return s.compare(p, strlen(t), t) == 0; }
    void parseChildren(XmlNode &parent, const std::string &closing) {
        for (;;) {
            // skip text
            while (p < s.size() && s[p] != '<') p++;
            if (p >= s.size()) {
                if (!closing.empty()) throw std::runtime_error("xml: unexpected end, expected </" + closing + ">");
                return;
            }
            if (startsWith("<!--")) {
                size_t e = s.find("-->", p);
                if (e == std::string::npos) throw std::runtime_error("xml: unterminated comment");
                p = e + 3;
                continue;
            }
            if (startsWith("<?")) {
                size_t e = s.find("?>", p);
                if (e == std::string::npos) throw std::runtime_error("xml: unterminated declaration");
                p = e + 2;
                continue;
            }
            if (startsWith("</")) {
                size_t e = s.find('>', p);
                std::string nm = s.substr(p + 2, e - p - 2);
                while (!nm.empty() && isspace((unsigned char)nm.back())) nm.pop_back();
                if (nm != closing) throw std::runtime_error("xml: mismatched </" + nm + ">, expected </" + closing + ">");
                p = e + 1;
                return;
            }
            // element
            p++;
            auto node = std::make_shared<XmlNode>();
            while (p < s.size() && !isspace((unsigned char)s[p]) && s[p] != '>' && s[p] != '/') node->name.push_back(s[p++]);
            bool selfClose = false;
            for (;;) {
                skipWs();
                if (p >= s.size()) throw std::runtime_error("xml: unterminated tag");
                if (s[p] == '/') { selfClose = true; p++; continue; }
                if (s[p] == '>') { p++; break; }
                std::string k;
                while (p < s.size() && s[p] != '=' && !isspace((unsigned char)s[p])) k.push_back(s[p++]);
                skipWs();
                if (s[p] != '=') throw std::runtime_error("xml: expected '=' after attribute " + k);
                p++;
                skipWs();
                const char q = s[p++];
                if (q != '"' && q != '\'') throw std::runtime_error("xml: expected quote");
                std::string v;
                while (p < s.size() && s[p] != q) v.push_back(s[p++]);
                p++;
                node->attrs.push_back({k, v});
            }
            if (!selfClose) parseChildren(*node, node->name);
            parent.children.push_back(node);
        }
    }
};

}  // namespace lmc_host
