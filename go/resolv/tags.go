package resolv

// Tags is a bitmask of up to 64 tags (tags.go:8). Has is "any bit in common" (tags.go:29-31); Unset is an XOR in the
// reference (tags.go:18-20) and stays one here.
type Tags uint64

var nextTagBit uint

// NewTag hands out the next free bit (tags.go:73-83 keeps names too; the name is accepted and ignored here).
func NewTag(name string) Tags {
	t := Tags(1) << nextTagBit
	nextTagBit++
	return t
}

func (t *Tags) Set(v Tags)     { *t |= v }
func (t *Tags) Unset(v Tags)   { *t ^= v }
func (t *Tags) Clear()         { *t = 0 }
func (t Tags) Has(v Tags) bool { return t&v > 0 }
func (t Tags) IsEmpty() bool   { return t == 0 }
